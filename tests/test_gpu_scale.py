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
