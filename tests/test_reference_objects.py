"""Reference-shaped objects (rebuilt from dumps of the real pyLatticeDSO object graph, tests/golden/objgraph_*.npz)
through the host layer.  CPU part: flattening + boundary-condition extraction of the drop-in reproduce, with the
oracle as solver, exactly what the reference's own write-back left on its objects.  GPU part (test_gpu_dropin.py):
the same comparison with the CUDA path as solver."""
import numpy as np
import pytest

from conftest import E_MOD, NU, load_golden
from fake_lattice import lattice_from_dump
from oracle import lattice_oracle as orc

CASES = ["bcc322_pen", "octet223_graded"]


@pytest.mark.parametrize("case", CASES)
def test_flatten_and_bcs_of_dumped_reference_objects(case):
    from pylatticedso_b200 import mesh as M
    G = load_golden(f"objgraph_{case}.npz")
    lat = lattice_from_dump(G)
    mesh = M.flatten_lattice(lat, None, "gmsh")
    assert mesh.n_points == len(G["p_index"]) and np.array_equal(mesh.point_index, G["p_index"])
    if case == "bcc322_pen":      # joint penalisation: 1.5 x radius on the beam_mod segments, flagged in the dump
        assert G["b_mod"].any() and set(np.unique(mesh.chain)) == {1.0, 1.5}
    fixed, g, f = M.bc_arrays_from_lattice(lat, mesh)
    assert int(fixed.sum()) == int(G["p_fixed"].sum())
    K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
    u, R = orc.solve_static(K, fixed.astype(bool), g, f)
    un = u.reshape(-1, 6)[: mesh.n_points]
    ref = G["u_points_expected"]
    assert np.abs(un - ref).max() <= 1e-10 * np.abs(ref).max()
    # the reference's k-fold accumulation of reactions on clamped nodes shared by k cells
    Rn = R.reshape(-1, 6)[: mesh.n_points]
    k_fold = np.zeros(mesh.n_points)
    pos = {int(i): k for k, i in enumerate(G["p_index"])}
    for c in lat.cells:
        for p in c.points_cell:
            if 1 in p.fixed_DOF:
                k_fold[pos[p.index]] += 1
    assert k_fold.max() > 1
    assert np.abs(Rn * k_fold[:, None] - G["reaction_points_expected"]).max() <= 1e-9 * np.abs(G["reaction_points_expected"]).max()
    # get_global_displacement: free DOFs of cell-boundary nodes, in the dumped (node, DOF) order
    for k, p in enumerate(mesh.meta["points"]):
        p.displacement_vector[:] = [float(v) for v in un[k]]
    xsol, idx = lat.get_global_displacement()
    assert xsol.shape == G["xsol_expected"].shape
    assert np.abs(xsol - G["xsol_expected"]).max() <= 1e-10 * np.abs(G["xsol_expected"]).max()
    assert np.array_equal(np.asarray(idx), G["global_displacement_index"])
