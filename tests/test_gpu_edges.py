"""GPU: edge cases -- tiny / ragged meshes, unconnected nodes, zero right-hand side, zero iterations,
cells without interior nodes."""
import numpy as np
import pytest

from conftest import E_MOD, NU
from oracle import lattice_oracle as orc

pytestmark = pytest.mark.gpu


def t(ctx, a, d):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)


def test_single_element_and_unconnected_node(ctx):
    xyz = np.array([[0.0, 0, 0], [0.3, 0.4, 1.2], [5.0, 5, 5]])          # node 2 has no element
    en0, en1, rad = np.array([0], np.int32), np.array([1], np.int32), np.array([0.07])
    rowptr, colidx = ctx.bsr_pattern(t(ctx, en0, np.int32), t(ctx, en1, np.int32), 3)
    assert rowptr.cpu().tolist() == [0, 2, 4, 4] and colidx.cpu().tolist() == [0, 1, 0, 1]
    vals = ctx.assemble_bsr(t(ctx, xyz[:, 0], np.float64), t(ctx, xyz[:, 1], np.float64), t(ctx, xyz[:, 2], np.float64),
                            t(ctx, en0, np.int32), t(ctx, en1, np.int32), t(ctx, rad, np.float64), 3, 4, E_MOD, NU)
    Ke = orc.element_stiffness(xyz[[0]], xyz[[1]], rad, E_MOD, NU)[0]
    got = vals.cpu().numpy().reshape(4, 6, 6)
    ref = [Ke[:6, :6], Ke[:6, 6:], Ke[6:, :6], Ke[6:, 6:]]
    for a, b in zip(got, ref):
        assert np.abs(a - b).max() < 1e-12 * np.abs(Ke).max()
    # clamp node 0 and the floating node, pull node 1: the cantilever solves, the unconnected node stays at 0
    fixed = np.zeros(18, np.uint8); fixed[:6] = 1; fixed[12:] = 1
    f = np.zeros(18); f[6 + 2] = 1e-3
    vbc, b = ctx.apply_dirichlet(rowptr, colidx, vals, t(ctx, fixed, np.uint8), t(ctx, np.zeros(18), np.float64), t(ctx, f, np.float64))
    u, info = ctx.pcg(rowptr, colidx, vbc, b, tol=1e-12, maxiter=200, precond=2)
    K = orc.assemble_csr(xyz, np.array([[0, 1]]), rad, E_MOD, NU)
    uo, _ = orc.solve_static(K[:12][:, :12], fixed[:12].astype(bool), np.zeros(12), f[:12])
    assert info["info"] in (0, 5) and np.abs(u.cpu().numpy()[:12] - uo).max() < 1e-9 * np.abs(uo).max()
    assert float(u[12:].abs().max()) == 0.0


@pytest.mark.parametrize("n_nodes", [1, 39, 40, 41, 81])
def test_ragged_row_counts(ctx, n_nodes):
    """Row counts around the 5-rows-per-warp / 40-rows-per-CTA tiling of the solver kernels."""
    rng = np.random.default_rng(n_nodes)
    xyz = rng.standard_normal((n_nodes + 1, 3))
    en0 = np.arange(n_nodes, dtype=np.int32); en1 = en0 + 1          # a chain of n_nodes elements
    rad = rng.uniform(0.02, 0.05, n_nodes)
    N = n_nodes + 1
    rowptr, colidx = ctx.bsr_pattern(t(ctx, en0, np.int32), t(ctx, en1, np.int32), N)
    vals = ctx.assemble_bsr(t(ctx, xyz[:, 0], np.float64), t(ctx, xyz[:, 1], np.float64), t(ctx, xyz[:, 2], np.float64),
                            t(ctx, en0, np.int32), t(ctx, en1, np.int32), t(ctx, rad, np.float64), N, colidx.numel(), E_MOD, NU)
    K = orc.assemble_csr(xyz, np.stack([en0, en1], 1), rad, E_MOD, NU)
    v = rng.standard_normal(6 * N)
    y = ctx.spmv(rowptr, colidx, vals, t(ctx, v, np.float64)).cpu().numpy()
    assert np.abs(y - K @ v).max() < 1e-12 * np.abs(K @ v).max()
    fixed = np.zeros(6 * N, np.uint8); fixed[:6] = 1
    f = np.zeros(6 * N); f[-4] = 1e-4
    vbc, b = ctx.apply_dirichlet(rowptr, colidx, vals, t(ctx, fixed, np.uint8), t(ctx, np.zeros(6 * N), np.float64), t(ctx, f, np.float64))
    u, info = ctx.pcg(rowptr, colidx, vbc, b, tol=1e-10, maxiter=100000, precond=2)
    uo, _ = orc.solve_static(K, fixed.astype(bool), np.zeros(6 * N), f)
    assert info["info"] in (0, 5) and np.abs(u.cpu().numpy() - uo).max() < 1e-6 * np.abs(uo).max()


def test_zero_rhs_and_zero_iterations(ctx):
    from pylatticedso_b200 import mesh as M
    m = M.mesh_from_synthetic(M.synthetic_lattice("BCC", (2, 2, 2), [0.05]), 1)
    fixed, g, f = M.compression_bc(m, value=0.0)                        # nothing imposed, nothing loaded
    dev = lambda a, d: t(ctx, a, d)
    rowptr, colidx = ctx.bsr_pattern(dev(m.en0, np.int32), dev(m.en1, np.int32), m.n_nodes)
    vals = ctx.assemble_bsr(dev(m.x, np.float64), dev(m.y, np.float64), dev(m.z, np.float64), dev(m.en0, np.int32),
                            dev(m.en1, np.int32), dev(m.rad, np.float64), m.n_nodes, colidx.numel(), E_MOD, NU)
    vbc, b = ctx.apply_dirichlet(rowptr, colidx, vals, dev(fixed, np.uint8), dev(g, np.float64), dev(f, np.float64))
    for classic in (False, True):
        u, info = ctx.pcg(rowptr, colidx, vbc, b, tol=1e-8, maxiter=50, precond=1, classic=classic)
        assert info["info"] == 0 and info["iters"] == 0 and float(u.abs().max()) == 0.0
    b2 = b.clone(); b2[7] = 1.0
    u, info = ctx.pcg(rowptr, colidx, vbc, b2, tol=1e-8, maxiter=0, precond=1)
    assert info["info"] == 1 and info["iters"] == 0 and float(u.abs().max()) == 0.0     # conjugate_gradient_solver.py:73


def test_schur_of_a_cell_without_interior_nodes(ctx):
    """Octet cell at one element per strut: all 14 joints lie on the cell boundary, n_I = 0, S = K_cell
    (SURVEY.md section 8a, row A7)."""
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.schur import bcc_cell_order_nodes, local_cell_mesh
    lat = M.synthetic_lattice("Octet", (1, 1, 1), [0.04])
    mesh = M.mesh_from_synthetic(lat, 1)
    bnd = bcc_cell_order_nodes(lat.pxyz, (0, 1, 0, 1, 0, 1))
    assert len(bnd) == 14 == mesh.n_nodes
    perm, xyz, l0, l1 = local_cell_mesh(mesh, bnd)
    S = ctx.schur_batch(t(ctx, xyz[None], np.float64), t(ctx, l0, np.int32), t(ctx, l1, np.int32),
                        t(ctx, mesh.rad[None], np.float64), 14, E_MOD, NU)[0].cpu().numpy()
    K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU).toarray()
    d = (bnd[:, None] * 6 + np.arange(6)[None, :]).ravel()
    assert S.shape == (84, 84) and np.abs(S - K[np.ix_(d, d)]).max() < 1e-12 * np.abs(K).max()
