"""GPU: size-independent properties at BASELINE.json's full sizes (no oracle solve at that size)."""
import numpy as np
import pytest

from conftest import E_MOD, NU

pytestmark = pytest.mark.gpu


def test_bcc20_m2_compression_full_size(ctx):
    """configs[1]: BCC 20^3, m = 2 -> 81 261 nodes / 128 000 elements / 487 566 DOF."""
    import torch
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.fem import BeamFEM
    lat = M.synthetic_lattice("BCC", (20, 20, 20), [0.05])
    m = M.mesh_from_synthetic(lat, 2)
    assert (m.n_nodes, m.n_elems, m.n_dof) == (81261, 128000, 487566)
    fixed, g, f = M.compression_bc(m)
    fem = BeamFEM(m, E_MOD, NU, ctx=ctx)
    u, R, info = fem.solve(fixed, g, f, tol=1e-8, maxiter=200000, precond=2)
    assert info["info"] == 0 and info["relres"] <= 1e-8
    assert fem.nnzb == m.n_nodes + 2 * m.n_elems
    # true residual of the constrained system, recomputed with the plain SpMV
    fx = torch.from_numpy(fixed).to(ctx.device).bool()
    r = R - torch.from_numpy(f).to(ctx.device)
    assert float(r[~fx].norm()) <= 2e-8 * info["norm_b"]     # true residual vs the recurrence's 1e-8 |b|
    assert float((u[fx] - torch.from_numpy(g).to(ctx.device)[fx]).abs().max()) < 1e-14
    # global equilibrium: the reactions balance (no external load)
    Rn = R.reshape(-1, 6)
    assert float(Rn[:, :3].sum(0).abs().max()) < 1e-6 * float(Rn[:, 2].abs().sum())
    # compression: top plate pushed down -> negative vertical reaction on top, positive at the bottom
    top = torch.from_numpy(M.surface_nodes(lat.pxyz, "Zmax")).to(ctx.device)
    assert float(Rn[top, 2].sum()) < 0
    # symmetry of the lattice under x <-> y swap shows up in the displacement field
    un = u.reshape(-1, 6)[: m.n_points].cpu().numpy()
    key = {tuple(np.round(p, 6)): k for k, p in enumerate(lat.pxyz)}
    idx = np.array([key[(p[1], p[0], p[2])] for p in np.round(lat.pxyz, 6)])
    assert np.abs(un[:, 2] - un[idx, 2]).max() < 5e-6 * np.abs(un[:, 2]).max()   # a tol = 1e-8 solve: error ~ cond * 1e-8
    assert np.abs(un[:, 0] - un[idx, 1]).max() < 5e-6 * np.abs(un[:, 2]).max()


def test_octet40_graded_gradient_full_size(ctx):
    """configs[2]: Octet 40^3 graded radii, 1 555 200 elements; gradient properties:
    linearity of the adjoint form and consistency with a directional finite difference of the compliance."""
    import torch
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200 import lib as L
    lat = M.synthetic_lattice("Octet", (40, 40, 40), [0.03], grad_radius=("linear", [False, False, True], [0, 0, 0.0125]))
    m = M.mesh_from_synthetic(lat, 1)
    assert (m.n_nodes, m.n_elems, m.n_dof) == (265721, 1555200, 1594326)
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)
    x, y, z, en0, en1, rad = t(m.x, np.float64), t(m.y, np.float64), t(m.z, np.float64), t(m.en0, np.int32), t(m.en1, np.int32), t(m.rad, np.float64)
    rowptr, colidx = ctx.bsr_pattern(en0, en1, m.n_nodes)
    assert colidx.numel() == m.n_nodes + 2 * m.n_elems
    vals = ctx.assemble_bsr(x, y, z, en0, en1, rad, m.n_nodes, colidx.numel(), E_MOD, NU)
    rng = np.random.default_rng(0)
    u = t(rng.standard_normal(m.n_dof) * 1e-3, np.float64)
    group = t(m.cell_of_elem, np.int32)
    ng = 64000
    g = ctx.compliance_grad(x, y, z, en0, en1, rad, group, ng, u, E_MOD, NU)
    # sum_p g_p * dr_p  ==  -u^T (dK/dr . dr) u  with dK assembled by the drad path (chain = dr of the element)
    dr = rng.uniform(-1, 1, ng)
    chain = t(dr[m.cell_of_elem], np.float64)
    dvals = ctx.assemble_bsr(x, y, z, en0, en1, rad, m.n_nodes, colidx.numel(), E_MOD, NU, drad=True, chain=chain)
    dKu = ctx.spmv(rowptr, colidx, dvals, u)
    lhs = float((g * t(dr, np.float64)).sum())
    rhs = -float((u * dKu).sum())
    assert abs(lhs - rhs) < 1e-9 * abs(rhs)
    # and against a central difference of u^T K(r) u along dr
    h = 1e-6
    vp = ctx.assemble_bsr(x, y, z, en0, en1, rad + h * chain, m.n_nodes, colidx.numel(), E_MOD, NU)
    ep = float((u * ctx.spmv(rowptr, colidx, vp, u)).sum())
    vm = ctx.assemble_bsr(x, y, z, en0, en1, rad - h * chain, m.n_nodes, colidx.numel(), E_MOD, NU)
    em = float((u * ctx.spmv(rowptr, colidx, vm, u)).sum())
    assert abs(-(ep - em) / (2 * h) - lhs) < 1e-6 * abs(lhs)
    # atomic and gather assembly agree
    va = ctx.assemble_bsr(x, y, z, en0, en1, rad, m.n_nodes, colidx.numel(), E_MOD, NU, mode=L.ASM_ATOMIC)
    assert float((va - vals).abs().max()) < 1e-12 * float(vals.abs().max())


def test_config0_bcc5_parity_mode_against_oracle(ctx):
    """BASELINE configs[0]: BCC 5^3 with joint penalisation on the reference's gmsh subdivision
    (18 333 nodes / 18 992 elements / 109 998 DOF, mesh and BCs dumped from the reference object graph):
    displacements and reactions at the lattice points against the oracle's direct solve, 1e-8."""
    from conftest import load_golden, mesh_from_npz
    from pylatticedso_b200.fem import BeamFEM
    G = load_golden("c1_parity_bcc555.npz")
    m = mesh_from_npz(G)
    assert (m.n_nodes, m.n_elems, m.n_dof) == (18333, 18992, 109998)
    n = int(G["n_dof"])
    fixed = np.unpackbits(G["fixed"])[:n].astype(np.uint8)
    assert int(fixed.sum()) == 252
    g = np.zeros(n); g[G["g_nonzero_idx"]] = G["g_nonzero_val"]
    fem = BeamFEM(m, E_MOD, NU, ctx=ctx)
    u, R, info = fem.solve(fixed, g, np.zeros(n), tol=1e-13, maxiter=2000000, precond=2)
    assert info["info"] in (0, 5)
    up = u.cpu().numpy().reshape(-1, 6)[: m.n_points]
    Rp = R.cpu().numpy().reshape(-1, 6)[: m.n_points]
    uo, Ro = G["u_points_oracle"], G["reactions_points_oracle"]
    assert np.abs(up - uo).max() < 1e-8 * np.abs(uo).max()
    fx = fixed.reshape(-1, 6)[: m.n_points].astype(bool)
    assert np.abs(Rp[fx] - Ro[fx]).max() < 1e-8 * np.abs(Ro[fx]).max()


def test_octet40_gradient_and_product_against_the_oracle_on_samples(ctx):
    """configs[2] at full size, oracle-checked on samples: the per-cell compliance gradient of 256 random cells and
    512 random rows of K u against the numpy oracle evaluated on exactly the elements involved."""
    import torch
    from oracle import lattice_oracle as orc
    from pylatticedso_b200 import mesh as M
    lat = M.synthetic_lattice("Octet", (40, 40, 40), [0.03], grad_radius=("linear", [False, False, True], [0, 0, 0.0125]))
    m = M.mesh_from_synthetic(lat, 1)
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)
    x, y, z, en0, en1, rad = t(m.x, np.float64), t(m.y, np.float64), t(m.z, np.float64), t(m.en0, np.int32), t(m.en1, np.int32), t(m.rad, np.float64)
    rng = np.random.default_rng(5)
    u = rng.standard_normal(m.n_dof) * 1e-3
    ng = 64000
    g = ctx.compliance_grad(x, y, z, en0, en1, rad, t(m.cell_of_elem, np.int32), ng, t(u, np.float64), E_MOD, NU).cpu().numpy()
    cells = rng.choice(ng, 256, replace=False)
    sel = np.flatnonzero(np.isin(m.cell_of_elem, cells))
    remap = -np.ones(ng, dtype=np.int64); remap[cells] = np.arange(cells.size)
    en = np.stack([m.en0[sel], m.en1[sel]], 1).astype(np.int64)
    go = orc.compliance_gradient(m.xyz, en, m.rad[sel], u, remap[m.cell_of_elem[sel]], cells.size, E_MOD, NU)
    assert np.abs(g[cells] - go).max() < 1e-10 * np.abs(go).max()
    # K u on sampled rows: the oracle assembles only the elements that touch the sampled nodes
    rowptr, colidx = ctx.bsr_pattern(en0, en1, m.n_nodes)
    vals = ctx.assemble_bsr(x, y, z, en0, en1, rad, m.n_nodes, colidx.numel(), E_MOD, NU)
    Ku = ctx.spmv(rowptr, colidx, vals, t(u, np.float64)).cpu().numpy().reshape(-1, 6)
    nodes = rng.choice(m.n_nodes, 512, replace=False)
    touch = np.flatnonzero(np.isin(m.en0, nodes) | np.isin(m.en1, nodes))
    en = np.stack([m.en0[touch], m.en1[touch]], 1).astype(np.int64)
    Ke = orc.element_stiffness(m.xyz[en[:, 0]], m.xyz[en[:, 1]], m.rad[touch], E_MOD, NU)
    dofs = (en[:, :, None] * 6 + np.arange(6)[None, None, :]).reshape(-1, 12)
    fe = np.einsum("eab,eb->ea", Ke, u[dofs])
    ref = np.zeros((m.n_nodes, 6))
    np.add.at(ref, en[:, 0], fe[:, :6])
    np.add.at(ref, en[:, 1], fe[:, 6:])
    assert np.abs(Ku[nodes] - ref[nodes]).max() < 1e-11 * np.abs(ref[nodes]).max()


def test_config3_bcc60_schur_batch_full_size_against_the_oracle_on_samples(ctx):
    """configs[3] at full size and at the reference's mesh density: 216 000 BCC cells (18 elements per strut, 870 DOF
    -> 48 boundary DOF each) with random radii through the star-cell kernel, values and dS/dr; 12 sampled cells
    against the oracle's DENSE condensation and a central difference of it; symmetry of every matrix."""
    import torch
    from oracle import lattice_oracle as orc
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.schur import bcc_cell_order_nodes, synthetic_cell_batch
    rng = np.random.default_rng(11)
    n_cells = 216000
    radii = rng.uniform(0.02, 0.08, n_cells)
    batch, bnd = synthetic_cell_batch(ctx, "BCC", radii, 18, E_MOD, NU, with_gradients=True)
    S, dS = batch.schur(with_gradients=True)
    assert tuple(S.shape) == (n_cells, 48, 48) and tuple(dS.shape)[:1] == (n_cells,)
    assert float((S - S.transpose(1, 2)).abs().max()) < 1e-10 * float(S.abs().max())
    lat = M.synthetic_lattice("BCC", (1, 1, 1), [1.0])
    mesh = M.mesh_from_synthetic(lat, 18)
    en = np.stack([mesh.en0, mesh.en1], 1).astype(np.int64)
    bdofs = (np.asarray(bnd)[:, None] * 6 + np.arange(6)[None, :]).ravel()
    pick = rng.choice(n_cells, 12, replace=False)

    def dense(r):
        return orc.schur_complement(orc.assemble_csr(mesh.xyz, en, np.full(mesh.n_elems, r), E_MOD, NU), bdofs)
    for c in pick:
        So = dense(radii[c])
        assert np.abs(S[c].cpu().numpy() - So).max() < 1e-10 * np.abs(So).max()
    c = pick[0]
    h = 1e-6
    fd = (dense(radii[c] + h) - dense(radii[c] - h)) / (2 * h)
    got = dS[c].cpu().numpy().reshape(48, 48)
    assert np.abs(got - fd).max() < 1e-6 * np.abs(fd).max()


def test_config4_octet100_full_size_solve(ctx):
    """configs[4] at full size on one GPU (24 361 806 DOF): the matrix-free solve to 1e-8, its true residual through the
    independent un-eliminated product, 512 sampled rows of K u against the numpy oracle's element matrices, and the
    assembled product (15 GB matrix) against the matrix-free one on the solution."""
    import torch
    from oracle import lattice_oracle as orc
    from pylatticedso_b200 import distributed as D, lib as L
    from pylatticedso_b200.fem import BeamFEM
    lm, _ = D.generate_slab("Octet", (100, 100, 100), [0.03], 1, 0, 1)
    assert (lm.n_nodes, lm.n_elems, lm.n_dof) == (4060301, 24120000, 24361806)
    fixed, g, f = D.compression_bc_local(lm)
    fem = BeamFEM(lm, E_MOD, NU, ctx=ctx)
    fem.build_pattern()
    assert fem.nnzb == lm.n_nodes + 2 * lm.n_elems
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)
    bc = (t(fixed, np.uint8), t(g, np.float64), t(f, np.float64))
    u, R, info = fem.solve_matrix_free(*bc, tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6)
    assert info["info"] == 0 and info["relres"] <= 1e-8 and info["true_relres"] <= 2e-8
    fx = bc[0].bool()
    assert float((u[fx] - bc[1][fx]).abs().max()) < 1e-14
    assert float((R - bc[2])[~fx].norm()) <= 3e-8 * info["norm_b"]          # R = K u from the un-eliminated operator
    Rn = R.reshape(-1, 6)
    assert float(Rn[:, :3].sum(0).abs().max()) < 1e-6 * float(Rn[:, 2].abs().sum())   # the reactions balance
    # sampled rows of K u against the oracle (elements that touch the sampled nodes only)
    rng = np.random.default_rng(9)
    uh = u.cpu().numpy()
    nodes = rng.choice(lm.n_nodes, 512, replace=False)
    touch = np.flatnonzero(np.isin(lm.en0, nodes) | np.isin(lm.en1, nodes))
    en = np.stack([lm.en0[touch], lm.en1[touch]], 1).astype(np.int64)
    Ke = orc.element_stiffness(lm.xyz[en[:, 0]], lm.xyz[en[:, 1]], lm.rad[touch], E_MOD, NU)
    dofs = (en[:, :, None] * 6 + np.arange(6)[None, None, :]).reshape(-1, 12)
    fe = np.einsum("eab,eb->ea", Ke, uh[dofs])
    ref = np.zeros((lm.n_nodes, 6))
    np.add.at(ref, en[:, 0], fe[:, :6])
    np.add.at(ref, en[:, 1], fe[:, 6:])
    Rh = Rn.cpu().numpy()
    scale = np.abs(fe).max()                                                  # rows of an equilibrium state cancel
    assert np.abs(Rh[nodes] - ref[nodes]).max() < 1e-11 * scale
    # assembled operator == matrix-free operator on the solution
    fem.assemble()
    Ra = ctx.spmv(fem.rowptr, fem.colidx, fem.vals, u)
    assert float((Ra - R).abs().max()) < 1e-11 * scale
