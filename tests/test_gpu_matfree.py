"""GPU: the matrix-free operator (csrc/matfree.cuh) against the assembled BSR path and the oracle.

Tolerances: a single product agrees with the assembled SpMV to 1e-13 relative (same coefficients, different
summation order); a solve to tol 1e-10 agrees with the oracle's sparse-direct displacement to 1e-8 relative
(north_star: displacements and reactions within 1e-8)."""
import numpy as np
import pytest

from conftest import E_MOD, NU

pytestmark = pytest.mark.gpu


def _dev(ctx, m):
    import torch
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)
    return t, (t(m.x, np.float64), t(m.y, np.float64), t(m.z, np.float64), t(m.en0, np.int32), t(m.en1, np.int32),
               t(m.rad, np.float64))


@pytest.mark.parametrize("geom,cells,mseg", [("BCC", (3, 2, 2), 2), ("Octet", (3, 3, 2), 1), ("BCC", (2, 2, 2), 3)])
def test_matfree_product_matches_assembled(ctx, geom, cells, mseg):
    import torch
    from pylatticedso_b200 import mesh as M
    lat = M.synthetic_lattice(geom, cells, [0.04], grad_radius=("linear", [True, False, True], [0.01, 0, 0.005]))
    m = M.mesh_from_synthetic(lat, mseg)
    t, (x, y, z, en0, en1, rad) = _dev(ctx, m)
    rowptr, colidx = ctx.bsr_pattern(en0, en1, m.n_nodes)
    vals = ctx.assemble_bsr(x, y, z, en0, en1, rad, m.n_nodes, colidx.numel(), E_MOD, NU)
    rng = np.random.default_rng(1)
    u = t(rng.standard_normal(m.n_dof), np.float64)
    # raw stiffness
    ctx.matfree_setup(x, y, z, en0, en1, rad, m.n_nodes, E_MOD, NU, fixed=None)
    y_mf = ctx.matfree_apply(u, eliminated=False)
    y_as = ctx.spmv(rowptr, colidx, vals, u)
    assert float((y_mf - y_as).abs().max()) <= 1e-13 * float(y_as.abs().max())
    # Dirichlet-eliminated operator and lifted right-hand side
    fixed, g, f = M.compression_bc(m)
    f = f + rng.standard_normal(m.n_dof) * (fixed == 0)
    fd, gd, fv = t(fixed, np.uint8), t(g, np.float64), t(f, np.float64)
    vbc, b_as = ctx.apply_dirichlet(rowptr, colidx, vals, fd, gd, fv, inplace=False)
    ctx.matfree_setup(x, y, z, en0, en1, rad, m.n_nodes, E_MOD, NU, fixed=fd)
    y_mf = ctx.matfree_apply(u, eliminated=True)
    y_as = ctx.spmv(rowptr, colidx, vbc, u)
    assert float((y_mf - y_as).abs().max()) <= 1e-13 * float(y_as.abs().max())
    b_mf = ctx.matfree_rhs(gd, fv)
    assert float((b_mf - b_as).abs().max()) <= 1e-13 * float(b_as.abs().max())


@pytest.mark.parametrize("precond", [0, 1, 2])
def test_matfree_solve_matches_oracle(ctx, precond):
    from oracle import lattice_oracle as O
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.fem import BeamFEM
    lat = M.synthetic_lattice("BCC", (3, 3, 3), [0.05])
    m = M.mesh_from_synthetic(lat, 2)
    fixed, g, f = M.compression_bc(m)
    fem = BeamFEM(m, E_MOD, NU, ctx=ctx)
    u, R, info = fem.solve_matrix_free(fixed, g, f, tol=1e-11, maxiter=50000, precond=precond)
    assert info["info"] in (0, 5) and info["true_relres"] <= 1e-10
    K = O.assemble_csr(m.xyz, np.stack([m.en0, m.en1], 1), m.rad, E_MOD, NU)
    uo, Ro = O.solve_static(K, fixed.astype(bool), g, f)
    assert np.abs(u.cpu().numpy() - uo).max() <= 1e-8 * np.abs(uo).max()
    assert np.abs(R.cpu().numpy() - Ro).max() <= 1e-8 * np.abs(Ro).max()
    # the assembled path lands on the same solution and takes (almost) the same number of iterations
    u2, R2, info2 = fem.solve(fixed, g, f, tol=1e-11, maxiter=50000, precond=precond)
    assert float((u - u2).abs().max()) <= 1e-8 * float(u2.abs().max())
    assert abs(info["iters"] - info2["iters"]) <= max(3, info2["iters"] // 20)   # rounding-level differences shift CG by a few %


def test_matfree_full_size_bcc20(ctx):
    """configs[1] without a stored matrix: same iteration count (+-2 %) and solution as the assembled path."""
    import torch
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.fem import BeamFEM
    lat = M.synthetic_lattice("BCC", (20, 20, 20), [0.05])
    m = M.mesh_from_synthetic(lat, 2)
    fixed, g, f = M.compression_bc(m)
    fem = BeamFEM(m, E_MOD, NU, ctx=ctx)
    u, R, info = fem.solve_matrix_free(fixed, g, f, tol=1e-8, precond=2, profile_iters=32)
    assert info["info"] == 0 and info["true_relres"] <= 2e-8
    u2, R2, info2 = fem.solve(fixed, g, f, tol=1e-8, precond=2, profile_iters=32)
    assert abs(info["iters"] - info2["iters"]) <= info2["iters"] // 20 + 2
    # two iterates stopped at |r| <= 1e-8 |b| differ by cond(A) * 1e-8; compare in the residual of the OTHER operator:
    # the matrix-free solution satisfies the assembled system to the same 2e-8
    fx = torch.from_numpy(fixed).to(ctx.device).bool()
    fv = torch.from_numpy(f).to(ctx.device)
    r_cross = ctx.spmv(fem.rowptr, fem.colidx, fem.vals, u) - fv
    assert float(r_cross[~fx].norm()) <= 2e-8 * info2["norm_b"]
    assert float((u - u2).abs().max()) <= 1e-3 * float(u2.abs().max())
    # tightened solves agree to 1e-8 relative
    ut, Rt, it = fem.solve_matrix_free(fixed, g, f, tol=1e-13, precond=2)
    ut2, Rt2, it2 = fem.solve(fixed, g, f, tol=1e-13, precond=2)
    print("tight diff u", float((ut - ut2).abs().max() / ut2.abs().max()), "R", float((Rt - Rt2).abs().max() / Rt2.abs().max()))
    assert float((ut - ut2).abs().max()) <= 1e-8 * float(ut2.abs().max())
    assert float((Rt - Rt2).abs().max()) <= 1e-8 * float(Rt2.abs().max())
    print("matfree", info["iters"], info["solve_ms"], info["spmv_ms"], info["update_ms"], "assembled", info2["iters"], info2["solve_ms"], info2["spmv_ms"], info2["update_ms"])


def test_matfree_errors(ctx):
    import torch
    from pylatticedso_b200 import lib as L
    from pylatticedso_b200 import mesh as M
    c2 = L.Context()
    b = torch.zeros(12, dtype=torch.float64, device=c2.device)
    with pytest.raises(L.LatticeB200Error):
        c2.pcg_matfree(b)                      # no resident operator
    lat = M.synthetic_lattice("BCC", (1, 1, 1), [0.05])
    m = M.mesh_from_synthetic(lat, 1)
    t, (x, y, z, en0, en1, rad) = _dev(c2, m)
    with pytest.raises(L.LatticeB200Error):
        c2.matfree_setup(x, y, z, en0, en1, rad, m.n_nodes, E_MOD, NU)   # no resident pattern
