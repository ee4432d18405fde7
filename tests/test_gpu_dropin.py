"""GPU: the drop-in layer (flatten -> C ABI -> write-back onto the objects) on stand-in lattice objects."""
import numpy as np
import pytest

from conftest import E_MOD, NU
from fake_lattice import FakeLattice
from oracle import lattice_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("matrix_free", [False, True, "condensed"])
def test_solve_fem_dropin_writes_back_like_the_reference(ctx, matrix_free):
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.fem import solve_FEM_B200
    lat = FakeLattice("BCC", (3, 2, 2), 0.05)
    lat.compression()
    if matrix_free == "condensed":
        xsol, model = solve_FEM_B200(lat, elements_per_strut=2, tol=1e-12, ctx=ctx, condense_struts=True)
    else:
        xsol, model = solve_FEM_B200(lat, elements_per_strut=2, tol=1e-12, ctx=ctx, matrix_free=matrix_free)
    assert model.info["info"] in (0, 5)
    mesh = M.mesh_from_synthetic(lat.syn, 2)
    fixed, g, f = M.compression_bc(mesh)
    K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
    uo, Ro = orc.solve_static(K, fixed.astype(bool), g, f)
    uo, Ro = uo.reshape(-1, 6), Ro.reshape(-1, 6)
    for p in lat.points:
        assert np.abs(np.array(p.displacement_vector) - uo[p.index]).max() < 1e-8 * np.abs(uo).max()
        if any(p.fixed_DOF):
            k = sum(1 for c in lat.cells if p in c.points_cell)   # reactions accumulate once per containing cell
            assert np.abs(np.array(p.reaction_force_vector) - k * Ro[p.index]).max() < 1e-8 * np.abs(Ro).max() * k
    xs, idx = lat.get_global_displacement()
    assert np.array_equal(xsol, xs) and len(idx) == len(xs)


def test_get_schur_complement_dropin(ctx):
    from pylatticedso_b200.schur import get_schur_complement
    from pylatticedso_b200.mesh import flatten_lattice, cell_boundary_dofs
    lat = FakeLattice("BCC", (2, 1, 1), 0.04)
    with pytest.raises(ValueError):
        get_schur_complement(lat, None, elements_per_strut=3, ctx=ctx)      # utils_schur.py:35-36
    S = get_schur_complement(lat, 1, elements_per_strut=3, ctx=ctx)
    mesh = flatten_lattice(lat, 1, 3)
    bnd = cell_boundary_dofs(lat.cells[1], mesh)
    K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
    So = orc.schur_complement(K, bnd)
    assert S.shape == (48, 48) and np.abs(S - So).max() < 1e-11 * np.abs(So).max()


def test_solve_ddm_dropin_matches_fem(ctx):
    from pylatticedso_b200.ddm import solve_DDM_B200
    from pylatticedso_b200.fem import solve_FEM_B200
    from pylatticedso_b200.schur import get_schur_complement
    a = FakeLattice("BCC", (2, 2, 2), 0.05); a.compression()
    b = FakeLattice("BCC", (2, 2, 2), 0.05); b.compression()
    x_fem, _ = solve_FEM_B200(a, elements_per_strut=2, tol=1e-13, ctx=ctx)
    S = get_schur_complement(b, 0, elements_per_strut=2, ctx=ctx)            # identical cells share one matrix
    for c in b.cells:
        c.schur_complement = S
    x_ddm, info, idx, rhs = solve_DDM_B200(b, tol=1e-13, ctx=ctx)
    assert info in (0, 5) and x_ddm.shape == x_fem.shape
    # compare_FEM_DDM.py:37-38
    assert np.linalg.norm(x_fem - x_ddm) / np.linalg.norm(x_fem) < 1e-8


@pytest.mark.parametrize("name", ["well_default", "ill_clamp"])
def test_conjugate_gradient_solver_dropin(ctx, name):
    """Same call as the reference's conjugate_gradient_solver(A, b, M, maxiter, tol, mintol, restart_every,
    alpha_max) -> (x, info), on a device operator; outputs frozen from the reference's own solver."""
    import torch
    from conftest import load_golden
    from pylatticedso_b200.pcg import BsrOperator, Jacobi, conjugate_gradient_solver
    G = load_golden("pcg_reference.npz")
    A = G["A_well"] if name.startswith("well") else G["A_ill"]
    nb = A.shape[0] // 6
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)
    op = BsrOperator(ctx, t(np.arange(0, nb * nb + 1, nb), np.int32), t(np.tile(np.arange(nb), nb), np.int32),
                     t(A.reshape(nb, 6, nb, 6).transpose(0, 2, 1, 3).reshape(-1), np.float64))
    maxiter, tol, mintol, restart, amax = G[f"{name}_params"]
    seen = []
    x, info = conjugate_gradient_solver(op, G[f"{name}_b"], M=Jacobi if bool(G[f"{name}_jacobi"]) else None,
                                        maxiter=int(maxiter), tol=tol, mintol=mintol, restart_every=int(restart),
                                        alpha_max=amax, callback=seen.append)
    assert info == int(G[f"{name}_info"]) and len(seen) == 1
    assert np.abs(x - G[f"{name}_x"]).max() < 1e-9 * np.abs(G[f"{name}_x"]).max()
    assert np.abs(op @ x - A @ x).max() < 1e-12 * np.abs(A @ x).max()
    with pytest.raises(TypeError):
        conjugate_gradient_solver(A, G[f"{name}_b"])          # a host matrix is not accepted: no CPU fallback


def test_compliance_gradient_lattice_parameter_order(ctx):
    """Parameter order cell.index * n_geom + j (lattice_opti.py:758-761) and the value against the oracle."""
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.fem import compliance_gradient_lattice, solve_FEM_B200
    lat = FakeLattice("BCC", (3, 1, 1), 0.05)
    for p in lat.points:                          # clamp Xmin, tip load on Xmax (dedup: every load once)
        if p.x == 0.0:
            p.fixed_DOF = [True] * 6
        elif p.x == 3.0:
            p.applied_force[2] = -0.025
    xsol, model = solve_FEM_B200(lat, elements_per_strut=3, tol=1e-13, dedup_point_loads=True, ctx=ctx)
    g = compliance_gradient_lattice(lat, model)
    mesh = model.fem.mesh
    en = np.stack([mesh.en0, mesh.en1], 1)
    go = orc.compliance_gradient(mesh.xyz, en, mesh.rad, model.u.cpu().numpy(), lat.syn.b_cell[mesh.beam_of_elem], 3,
                                 E_MOD, NU, chain=mesh.chain)
    assert g.shape == (3,) and np.abs(g - go).max() < 1e-9 * np.abs(go).max()
    assert (g < 0).all() and abs(g[0]) > abs(g[1]) > abs(g[2])     # thicker struts near the clamp help most
