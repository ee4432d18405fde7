"""GPU: batched per-cell Schur complements and the DDM interface operator."""
import numpy as np
import pytest

from conftest import E_MOD, NU, load_golden, mesh_from_npz
from oracle import lattice_oracle as orc

pytestmark = pytest.mark.gpu


def t(ctx, a, d):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)


@pytest.mark.parametrize("geom,tol", [("BCC", 1e-11), ("Hybrid1", 1e-11), ("Hybrid4", 1e-11)])
def test_schur_batch_reproduces_reference_goldens(ctx, geom, tol):
    """The 30 Schur matrices stored by the reference (dolfinx/PETSc) through lat_schur_batch."""
    from pylatticedso_b200.schur import local_cell_mesh
    G = load_golden(f"schur_{geom}.npz")
    SM = G["schur_matrices"]
    for i in range(SM.shape[0]):
        m = mesh_from_npz(G, f"c{i}_")
        bnd_nodes = G[f"c{i}_bnd"][::6] // 6
        perm, xyz, l0, l1 = local_cell_mesh(m, bnd_nodes)
        S = ctx.schur_batch(t(ctx, xyz[None], np.float64), t(ctx, l0, np.int32), t(ctx, l1, np.int32),
                            t(ctx, m.rad[None], np.float64), len(bnd_nodes), E_MOD, NU)[0].cpu().numpy()
        err = np.abs(S - SM[i]).max() / np.abs(SM[i]).max()
        assert err < tol, (geom, i, err)
        assert np.array_equal(S, S.T)
        # the same golden through the strut pre-pass (lat_schur_batch_chains): chains condensed first, joints after
        from pylatticedso_b200.schur import _chains_to_device, strut_chains
        ch = strut_chains(xyz, l0, l1, len(bnd_nodes))
        assert ch is not None
        S2 = ctx.schur_batch_chains(t(ctx, xyz[None], np.float64), t(ctx, l0, np.int32), t(ctx, l1, np.int32),
                                    t(ctx, m.rad[None], np.float64), _chains_to_device(ch, ctx.device), len(bnd_nodes),
                                    E_MOD, NU)[0].cpu().numpy()
        err2 = np.abs(S2 - SM[i]).max() / np.abs(SM[i]).max()
        assert err2 < tol, (geom, i, "chains", err2)
        assert np.array_equal(S2, S2.T)
        # and through lat_schur_batch_struts: the penalised BCC cell is a star (8 struts of 18 elements with the x1.5
        # segments inside, one interior joint) and runs in the half-warp kernel; the hybrids fall through to the chains path
        from pylatticedso_b200.schur import is_star
        S3 = ctx.schur_batch_struts(t(ctx, xyz[None], np.float64), t(ctx, l0, np.int32), t(ctx, l1, np.int32),
                                    t(ctx, m.rad[None], np.float64), _chains_to_device(ch, ctx.device), len(bnd_nodes),
                                    E_MOD, NU)[0].cpu().numpy()
        assert is_star(ch, len(bnd_nodes)) == (geom == "BCC")
        err3 = np.abs(S3 - SM[i]).max() / np.abs(SM[i]).max()
        assert err3 < tol, (geom, i, "struts", err3)
        assert np.abs(S3 - S3.T).max() <= 1e-13 * np.abs(S3).max()


def test_schur_batch_many_cells_and_gradients(ctx):
    """Config-4 style batch: one BCC cell per radius, n_I = 6 (m=1) and n_I = 102-like (m=3) meshes,
    against the oracle; analytic dS/dr against a central difference of the oracle."""
    from pylatticedso_b200.schur import synthetic_cell_batch
    from pylatticedso_b200 import mesh as M
    rng = np.random.default_rng(44)
    radii = 0.02 + 0.06 * rng.random(300)
    for m_ in (1, 3):
        batch, bnd = synthetic_cell_batch(ctx, "BCC", radii, m_, E_MOD, NU, with_gradients=True)
        S, dS = batch.schur(with_gradients=True)
        S, dS = S.cpu().numpy(), dS.cpu().numpy()
        lat = M.synthetic_lattice("BCC", (1, 1, 1), [1.0])
        mesh = M.mesh_from_synthetic(lat, m_)
        en = np.stack([mesh.en0, mesh.en1], 1)
        bd = (bnd[:, None] * 6 + np.arange(6)[None, :]).ravel()
        for c in (0, 17, 299):
            f = lambda r: orc.schur_complement(orc.assemble_csr(mesh.xyz, en, np.full(mesh.n_elems, r), E_MOD, NU), bd)
            So = f(radii[c])
            assert np.abs(S[c] - So).max() < 1e-11 * np.abs(So).max()
            h = 1e-6
            fd = (f(radii[c] + h) - f(radii[c] - h)) / (2 * h)
            assert np.abs(dS[c, 0] - fd).max() < 1e-6 * np.abs(fd).max()
        # the star kernel (differentiated strut pre-pass) against the dense route over all interior DOFs
        assert batch.star
        Sd, dSd = batch.schur(with_gradients=True, use_chains=False)
        assert np.abs(S - Sd.cpu().numpy()).max() < 1e-12 * np.abs(S).max()
        assert np.abs(dS - dSd.cpu().numpy()).max() < 1e-9 * np.abs(dS).max()
        # S is symmetric positive semi-definite with exactly the 6 rigid-body modes in its null space
        w = np.linalg.eigvalsh(S[5])
        assert w.min() > -1e-9 * w.max() and (np.abs(w) < 1e-9 * w.max()).sum() == 6


def test_ddm_matvec_matches_assembled_interface_operator(ctx):
    """y = sum_c B_c S_c B_c^T x against a dense numpy restatement of lattice_sim.py:1180-1252."""
    rng = np.random.default_rng(3)
    n_cells, nb, n_free = 50, 48, 400
    S = rng.standard_normal((n_cells, nb, nb)); S = S + S.transpose(0, 2, 1)
    gidx = np.full((n_cells, nb), -1, dtype=np.int32)
    for c in range(n_cells):
        pick = rng.choice(n_free, size=40, replace=False)
        gidx[c, rng.choice(nb, size=40, replace=False)] = pick
    x = rng.standard_normal(n_free)
    ufix = rng.standard_normal((n_cells, nb)) * (gidx < 0)
    for uf in (None, ufix):
        y = ctx.ddm_matvec(t(ctx, S, np.float64), t(ctx, gidx, np.int32), t(ctx, x, np.float64), n_free,
                           None if uf is None else t(ctx, uf, np.float64)).cpu().numpy()
        ref = np.zeros(n_free)
        for c in range(n_cells):
            u = np.where(gidx[c] >= 0, x[np.maximum(gidx[c], 0)], 0.0 if uf is None else uf[c])
            r = S[c] @ u
            np.add.at(ref, gidx[c][gidx[c] >= 0], r[gidx[c] >= 0])
        assert np.abs(y - ref).max() < 1e-12 * np.abs(ref).max()
    # shared S for all cells (uniform lattice: one cached Schur matrix, lattice_sim.py:857-883)
    y = ctx.ddm_matvec(t(ctx, S[0], np.float64), t(ctx, gidx, np.int32), t(ctx, x, np.float64), n_free).cpu().numpy()
    ref = np.zeros(n_free)
    for c in range(n_cells):
        u = np.where(gidx[c] >= 0, x[np.maximum(gidx[c], 0)], 0.0)
        r = S[0] @ u
        np.add.at(ref, gidx[c][gidx[c] >= 0], r[gidx[c] >= 0])
    assert np.abs(y - ref).max() < 1e-12 * np.abs(ref).max()


def test_ddm_interface_solve_equals_full_fem(ctx):
    """Static condensation is exact: per-cell Schur (lat_schur_batch) -> assembled interface operator
    (lat_assemble_cells_bsr) -> PCG must reproduce the full-lattice FEM displacements on the cell corners,
    and the matrix-free operator (lat_ddm_matvec) must equal the assembled one."""
    import torch
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.ddm import InterfaceProblem
    from pylatticedso_b200.fem import BeamFEM
    from pylatticedso_b200.schur import bcc_cell_order_nodes, synthetic_cell_batch
    n = (3, 2, 2)
    lat = M.synthetic_lattice("BCC", n, [0.05])
    mesh = M.mesh_from_synthetic(lat, 3)
    fixed, g, f = M.compression_bc(mesh)
    f = f.copy()
    fem = BeamFEM(mesh, E_MOD, NU, ctx=ctx)
    u, R, info = fem.solve(fixed, g, f, tol=1e-13, maxiter=200000)
    u_pts = u.cpu().numpy().reshape(-1, 6)[: mesh.n_points]
    # interface = the cell corners; every cell is a translate of the unit cell
    key = {tuple(np.round(p, 9)): k for k, p in enumerate(lat.pxyz)}
    corner = np.array([k for k, p in enumerate(lat.pxyz) if np.allclose(p, np.round(p))])
    g2i = -np.ones(lat.pxyz.shape[0], dtype=np.int64); g2i[corner] = np.arange(corner.size)
    unit = M.synthetic_lattice("BCC", (1, 1, 1), [1.0])
    order = bcc_cell_order_nodes(unit.pxyz, (0, 1, 0, 1, 0, 1))
    cells = []
    for i in range(n[0]):
        for j in range(n[1]):
            for k in range(n[2]):
                cells.append([g2i[key[tuple(np.round(unit.pxyz[o] + np.array([i, j, k]), 9))]] for o in order])
    cell_nodes = np.array(cells, dtype=np.int32)
    batch, _ = synthetic_cell_batch(ctx, "BCC", np.array([0.05]), 3, E_MOD, NU)
    S = batch.schur()[0]
    prob = InterfaceProblem(ctx, cell_nodes, corner.size, S)
    dof = (corner[:, None] * 6 + np.arange(6)[None, :]).ravel()
    ui, Ri, infoi, b = prob.solve(fixed[dof], g[dof], f[dof], tol=1e-13)
    assert infoi["info"] in (0, 5)   # 5: tol = 1e-13 is below what the true residual can reach in FP64
    ref = u_pts[corner]
    assert np.abs(ui.cpu().numpy().reshape(-1, 6) - ref).max() < 1e-8 * np.abs(ref).max()
    # matrix-free operator == assembled operator on a random vector (all DOFs free)
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal(6 * corner.size)).to(ctx.device)
    gidx = torch.from_numpy((cell_nodes[:, :, None] * 6 + np.arange(6)[None, None, :]).reshape(len(cells), -1).astype(np.int32)).to(ctx.device)
    y1 = ctx.ddm_matvec(S, gidx, x)
    y2 = ctx.spmv(prob.rowptr, prob.colidx, prob.vals, x)
    assert float((y1 - y2).abs().max()) < 1e-11 * float(y2.abs().max())
    # the plan-driven gather assembly (InterfaceProblem.assemble) == the atomic scatter assembly, is bit-reproducible, and
    # follows new Schur matrices on the same pattern (per-cell S, one cell with an absent node)
    va = ctx.assemble_cells_bsr(S, prob.cell_nodes, prob.rowptr, prob.colidx)
    assert float((va - prob.vals).abs().max()) < 1e-13 * float(va.abs().max())
    Sc = S[None] * torch.from_numpy(1.0 + 0.1 * rng.random(len(cells))).to(ctx.device)[:, None, None]
    v1 = prob.assemble(Sc).clone()
    v2 = prob.assemble(Sc)
    assert torch.equal(v1, v2)
    va = ctx.assemble_cells_bsr(Sc, prob.cell_nodes, prob.rowptr, prob.colidx)
    assert float((va - v1).abs().max()) < 1e-13 * float(va.abs().max())
    cn2 = cell_nodes.copy(); cn2[1, 3] = -1
    prob2 = InterfaceProblem(ctx, cn2, corner.size, Sc)
    va = ctx.assemble_cells_bsr(Sc, prob2.cell_nodes, prob2.rowptr, prob2.colidx)
    assert float((va - prob2.vals).abs().max()) < 1e-13 * float(va.abs().max())


def test_schur_dataset_roundtrip_in_reference_schema(ctx, tmp_path):
    """A GPU batch written in the reference's npz schema reproduces the reference's own stored dataset."""
    from pylatticedso_b200.schur import local_cell_mesh, load_schur_dataset, save_schur_dataset
    G = load_golden("schur_Hybrid4.npz")
    S_all = []
    for i in range(10):
        m = mesh_from_npz(G, f"c{i}_")
        bnd_nodes = G[f"c{i}_bnd"][::6] // 6
        perm, xyz, l0, l1 = local_cell_mesh(m, bnd_nodes)
        S_all.append(ctx.schur_batch(t(ctx, xyz[None], np.float64), t(ctx, l0, np.int32), t(ctx, l1, np.int32),
                                     t(ctx, m.rad[None], np.float64), len(bnd_nodes), E_MOD, NU)[0])
    import torch
    p = save_schur_dataset(str(tmp_path / "Schur_complement_Hybrid4.npz"), G["radius_values"], torch.stack(S_all))
    d = load_schur_dataset(p)
    assert len(d) == 10
    for i, r in enumerate(G["radius_values"]):
        ref = G["schur_matrices"][i]
        assert np.abs(d[tuple(r)] - ref).max() < 1e-11 * np.abs(ref).max()


def test_ddm_regular_bcc_per_cell_radii_equals_full_fem(ctx):
    """configs[3] in small (BCC 8^3 = 512 cells, per-cell radii, 3 elements per strut): batched condensation ->
    assembled interface operator -> PCG reproduces the full FEM displacements and reactions at the cell corners."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("ddm_config3", os.path.join(os.path.dirname(__file__), "..", "tools", "ddm_config3.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    out = mod.run(ctx, 8, 3, tol=1e-12, verbose=False)
    assert out["ddm_info"] in (0, 5) and out["u_rel"] < 1e-8 and out["R_rel"] < 1e-8
    assert out["interface_dof"] == 6 * 9 ** 3 and out["fem_dof"] > 10 * out["interface_dof"]


@pytest.mark.parametrize("m_", [1, 3])
def test_octet_cells_without_interior_joint_direct_path(ctx, m_):
    """Octet cells (14 joints, all on the cell boundary, 84 boundary DOF): the strut path assembles S directly
    (k_schur_direct); equal to the dense condensation of all interior nodes and to the oracle."""
    from oracle import lattice_oracle as orc
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.schur import synthetic_cell_batch
    rng = np.random.default_rng(5)
    radii = rng.uniform(0.02, 0.06, 40)
    batch, bnd = synthetic_cell_batch(ctx, "Octet", radii, m_, E_MOD, NU)
    assert batch.chains is not None and batch.chains["n_joints"] == len(bnd) == 14 and not batch.star
    S = batch.schur().cpu().numpy()
    Sd = batch.schur(use_chains=False).cpu().numpy()
    assert S.shape == (40, 84, 84)
    assert np.abs(S - Sd).max() < 1e-12 * np.abs(Sd).max()
    cell = M.mesh_from_synthetic(M.synthetic_lattice("Octet", (1, 1, 1), [1.0]), m_)
    en = np.stack([cell.en0, cell.en1], 1).astype(np.int64)
    bdofs = (np.asarray(bnd)[:, None] * 6 + np.arange(6)[None, :]).ravel()
    K = orc.assemble_csr(cell.xyz, en, np.full(cell.n_elems, radii[7]), E_MOD, NU)
    So = orc.schur_complement(K, bdofs) if m_ > 1 else K.toarray()[np.ix_(bdofs, bdofs)]
    assert np.abs(S[7] - So).max() < 1e-11 * np.abs(So).max()


# strut tables of four more reference geometries (src/pyLatticeDesign/geometries/*.json: constant data)
_BCCZ = [[0.5, 0.5, 0.5, 1, 1, 1], [0, 0, 0, 0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 1, 1, 0], [0, 0, 1, 0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 0, 1, 0],
         [1, 0, 1, 0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 0, 1, 1], [1, 0, 0, 0.5, 0.5, 0.5], [0.5, 0.5, 0, 0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5, 0.5, 1]]
_HYBRID2 = [[0.5, 0, 0, 0.5, 0.5, 0.5], [1, 0, 0.5, 0.5, 0.5, 0.5], [0.5, 0, 1, 0.5, 0.5, 0.5], [0, 0, 0.5, 0.5, 0.5, 0.5], [0.5, 1, 0, 0.5, 0.5, 0.5],
            [1, 1, 0.5, 0.5, 0.5, 0.5], [0.5, 1, 1, 0.5, 0.5, 0.5], [0, 1, 0.5, 0.5, 0.5, 0.5], [0, 0.5, 0, 0.5, 0.5, 0.5], [0, 0.5, 1, 0.5, 0.5, 0.5],
            [1, 0.5, 0, 0.5, 0.5, 0.5], [1, 0.5, 1, 0.5, 0.5, 0.5]]
_CUBIC = [[0, 0, 0, 0, 0, 1], [1, 0, 0, 1, 0, 1], [0, 1, 0, 0, 1, 1], [1, 1, 0, 1, 1, 1], [0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0], [1, 1, 0, 0, 1, 0],
          [1, 1, 0, 1, 0, 0], [0, 0, 1, 1, 0, 1], [0, 0, 1, 0, 1, 1], [1, 1, 1, 0, 1, 1], [1, 1, 1, 1, 0, 1]]
_KELVIN = [[0.5, 0.25, 0, 0.25, 0.5, 0], [0.5, 0.25, 0, 0.75, 0.5, 0], [0.5, 0.75, 0, 0.25, 0.5, 0], [0.5, 0.75, 0, 0.75, 0.5, 0],
           [0.5, 0.25, 1, 0.25, 0.5, 1], [0.5, 0.25, 1, 0.75, 0.5, 1], [0.5, 0.75, 1, 0.25, 0.5, 1], [0.5, 0.75, 1, 0.75, 0.5, 1],
           [0.5, 0, 0.25, 0.25, 0, 0.5], [0.5, 0, 0.25, 0.75, 0, 0.5], [0.5, 0, 0.75, 0.25, 0, 0.5], [0.5, 0, 0.75, 0.75, 0, 0.5],
           [0.5, 1, 0.25, 0.25, 1, 0.5], [0.5, 1, 0.25, 0.75, 1, 0.5], [0.5, 1, 0.75, 0.25, 1, 0.5], [0.5, 1, 0.75, 0.75, 1, 0.5],
           [0, 0.5, 0.25, 0, 0.25, 0.5], [0, 0.5, 0.25, 0, 0.75, 0.5], [0, 0.5, 0.75, 0, 0.25, 0.5], [0, 0.5, 0.75, 0, 0.75, 0.5],
           [1, 0.5, 0.25, 1, 0.25, 0.5], [1, 0.5, 0.25, 1, 0.75, 0.5], [1, 0.5, 0.75, 1, 0.25, 0.5], [1, 0.5, 0.75, 1, 0.75, 0.5],
           [0.5, 0.25, 0, 0.5, 0, 0.25], [0.25, 0.5, 0, 0, 0.5, 0.25], [0.75, 0.5, 0, 1, 0.5, 0.25], [0.5, 0.75, 0, 0.5, 1, 0.25],
           [0.25, 0, 0.5, 0, 0.25, 0.5], [0.75, 0, 0.5, 1, 0.25, 0.5], [0.75, 1, 0.5, 1, 0.75, 0.5], [0.25, 1, 0.5, 0, 0.75, 0.5],
           [0.5, 0, 0.75, 0.5, 0.25, 1], [0, 0.5, 0.75, 0.25, 0.5, 1], [0.5, 1, 0.75, 0.5, 0.75, 1], [1, 0.5, 0.75, 0.75, 0.5, 1]]


@pytest.mark.parametrize("name,table,n_bnd,kind", [("BCCZ", _BCCZ, 10, "star"), ("Hybrid2", _HYBRID2, 12, "star"),
                                                   ("Cubic", _CUBIC, 8, "direct"), ("Kelvin", _KELVIN, 24, "direct")])
def test_more_reference_geometries_take_the_fast_schur_kernels(ctx, name, table, n_bnd, kind):
    """Star cells beyond BCC and cells without an interior joint beyond Octet (Kelvin: 24 boundary joints, 144 DOF):
    the strut path (k_schur_star / k_schur_direct) equals the dense condensation of every interior node."""
    from pylatticedso_b200.schur import synthetic_cell_batch
    rng = np.random.default_rng(len(table))
    batch, bnd = synthetic_cell_batch(ctx, np.asarray(table, dtype=float), rng.uniform(0.02, 0.05, 24), 2, E_MOD, NU)
    assert len(bnd) == n_bnd and batch.chains is not None
    assert batch.star == (kind == "star") and (batch.chains["n_joints"] == n_bnd) == (kind == "direct")
    S = batch.schur().cpu().numpy()
    Sd = batch.schur(use_chains=False).cpu().numpy()
    assert S.shape == (24, 6 * n_bnd, 6 * n_bnd)
    assert np.abs(S - Sd).max() < 1e-12 * np.abs(Sd).max()


def test_octet_sensitivities_through_the_direct_kernel(ctx):
    """dS/dr of cells without an interior joint: S is linear in the super-elements, so the direct assembly applied to
    the forward-mode derivative of the strut pre-pass is the analytic sensitivity.  Against the dense route (analytic,
    all interior nodes) and a central difference of the values."""
    from pylatticedso_b200.schur import synthetic_cell_batch
    rng = np.random.default_rng(8)
    radii = rng.uniform(0.02, 0.05, 20)
    batch, bnd = synthetic_cell_batch(ctx, "Octet", radii, 2, E_MOD, NU, with_gradients=True)
    assert batch.direct and batch.chain_group is not None
    S, dS = batch.schur(with_gradients=True)
    Sd, dSd = batch.schur(with_gradients=True, use_chains=False)
    assert tuple(dS.shape) == (20, 1, 84, 84)
    assert float((S - Sd).abs().max()) < 1e-12 * float(Sd.abs().max())
    assert float((dS - dSd).abs().max()) < 1e-11 * float(dSd.abs().max())
    h = 1e-6
    Sp = synthetic_cell_batch(ctx, "Octet", radii + h, 2, E_MOD, NU)[0].schur()
    Sm = synthetic_cell_batch(ctx, "Octet", radii - h, 2, E_MOD, NU)[0].schur()
    fd = (Sp - Sm) / (2 * h)
    assert float((dS[:, 0] - fd).abs().max()) < 1e-6 * float(fd.abs().max())
