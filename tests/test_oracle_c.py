"""CPU: the C/OpenMP restatement (oracle/oracle_c.c) against the numpy oracle that the goldens pin."""
import numpy as np
import scipy.sparse as sp

from conftest import E_MOD, NU, load_golden
from oracle import lattice_oracle as orc
from oracle import oracle_c as oc
from pylatticedso_b200 import mesh as M


def test_c_element_and_assembly_match_numpy_oracle():
    m = M.mesh_from_synthetic(M.synthetic_lattice("Octet", (2, 2, 2), [0.03]), 2)
    en = np.stack([m.en0, m.en1], 1)
    Ke = oc.elem_stiffness(m.xyz, en, m.rad, E_MOD, NU)
    ref = orc.element_stiffness(m.xyz[m.en0], m.xyz[m.en1], m.rad, E_MOD, NU)
    assert np.abs(Ke - ref).max() < 1e-13 * np.abs(ref).max()
    K = orc.assemble_csr(m.xyz, en, m.rad, E_MOD, NU)
    data = oc.assemble_csr_values(en, Ke, K.indptr, K.indices)
    assert np.abs(data - K.data).max() < 1e-12 * np.abs(K.data).max()


def test_c_pcg_is_the_reference_pcg():
    G = load_golden("pcg_reference.npz")
    for name in ("well_default", "well_ddm", "ill_clamp"):
        A = sp.csr_matrix(G["A_well"] if name.startswith("well") else G["A_ill"])
        maxiter, tol, mintol, restart, amax = G[f"{name}_params"]
        dinv = 1.0 / A.diagonal() if bool(G[f"{name}_jacobi"]) else None
        x, info, it = oc.pcg(A.indptr, A.indices, A.data, G[f"{name}_b"], dinv, int(maxiter), tol, mintol, int(restart), amax)
        assert info == int(G[f"{name}_info"]) and it == int(G[f"{name}_iters"])
        assert np.abs(x - G[f"{name}_x"]).max() < 1e-9 * np.abs(G[f"{name}_x"]).max()
    assert oc.num_threads() >= 1
