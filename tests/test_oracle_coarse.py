"""CPU: the two-level preconditioner restated in oracle/coarse_oracle.py -- it must not change the solution (checked
against the direct solve of oracle/lattice_oracle.py, which the reference fixtures pin) and it must cut the iterations
of the stretch-dominated octet lattice."""
import numpy as np

from conftest import E_MOD, NU
from oracle import coarse_oracle as co
from oracle import lattice_oracle as lo


def _system(geom, n, m_el):
    from pylatticedso_b200 import mesh as M
    lat = M.synthetic_lattice(geom, (n, n, n), [0.05])
    m = M.mesh_from_synthetic(lat, m_el)
    K = lo.assemble_csr(m.xyz, np.stack([m.en0, m.en1], 1), m.rad, E_MOD, NU)
    fixed, g, f = M.compression_bc(m)
    Kbc, rhs = lo.apply_dirichlet(K, fixed, g, f)
    return m, K, Kbc.tocsr(), rhs, fixed, g, f


def test_rigid_body_modes_have_no_strain_energy():
    """K Z = 0 for the unconstrained operator: every column of Z is a rigid-body motion of its aggregate, so the only
    energy comes from the struts that cross aggregate borders; with ONE aggregate the coarse matrix vanishes."""
    m, K, *_ = _system("BCC", 3, 1)
    Z = co.rigid_body_modes(m.xyz, np.zeros(m.n_nodes, dtype=np.int64), 1)
    assert np.abs(K @ Z.toarray()).max() < 1e-9 * np.abs(K.data).max()


def test_two_level_pcg_same_solution_fewer_iterations():
    m, K, Kbc, rhs, fixed, g, f = _system("Octet", 6, 1)
    u_ref = lo.solve_static(K, fixed, g, f)[0]
    Dinv = co.block_jacobi_inverse(Kbc, m.n_nodes)
    agg, n_agg = co.box_aggregates(m.xyz, 27)
    assert n_agg == 27
    Z = co.rigid_body_modes(m.xyz, agg, n_agg, fixed)
    E = co.coarse_matrix(Kbc, Z)
    assert np.abs(E - E.T).max() < 1e-12 * np.abs(E).max()
    np.testing.assert_allclose(E, co.coarse_matrix(K, Z), rtol=0, atol=1e-12 * np.abs(E).max())   # elimination does not matter
    Einv = co.coarse_pinv(E)
    x1, it1 = co.two_level_pcg(Kbc, rhs, Dinv, tol=1e-10)
    x2, it2 = co.two_level_pcg(Kbc, rhs, Dinv, Z, Einv, tol=1e-10)
    scale = np.abs(u_ref).max()
    assert np.abs(x1 - u_ref).max() < 1e-7 * scale and np.abs(x2 - u_ref).max() < 1e-7 * scale
    assert it2 < 0.8 * it1          # 68 -> 48 at this size; the gain grows with the lattice (228 -> 53 at 24^3)
