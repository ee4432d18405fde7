"""Domain-decomposition (static condensation) solve on the GPU: the device counterpart of
``LatticeSim.solve_DDM`` (lattice_sim.py:1111-1176).

    per-cell Schur complements (lat_schur_batch)            -- SchurComplement, schur_complement.py:75-147
 -> interface operator K_G = sum_c P_c^T S_c P_c as BSR     -- build_preconditioner, lattice_sim.py:1351-1415
 -> Dirichlet elimination + block-Jacobi PCG on K_G         -- replaces CG on the Python operator
                                                               calculate_reaction_force_global (:1180-1252)
 -> reactions R = K_G u                                     -- update_reaction_force_each_cell (:1204-1223)

The reference iterates on the free interface DOFs with a Python loop over all cells per CG iteration and
preconditions with a SuperLU factorisation of the same assembled matrix; here the assembled matrix itself is
solved.  ``lat_ddm_matvec`` (the matrix-free operator) remains available through ``Context.ddm_matvec``.
"""
from __future__ import annotations

import numpy as np

from . import lib as L
from .mesh import NDOF


def cell_pair_elements(cell_nodes):
    """All unordered node pairs of every cell as virtual 2-node elements (pattern input), pair-major: one strided
    column copy per pair (28 for eight corner nodes) instead of a fancy-indexed gather of every entry
    (6 M pairs at BASELINE config 3: 115 -> ~15 ms on the host)."""
    cn = np.ascontiguousarray(cell_nodes, dtype=np.int32)
    nc, nbn = cn.shape
    ia, ib = np.triu_indices(nbn, k=1)
    a = np.empty((ia.shape[0], nc), dtype=np.int32)
    b = np.empty((ia.shape[0], nc), dtype=np.int32)
    for p, (i, j) in enumerate(zip(ia, ib)):
        a[p] = cn[:, i]
        b[p] = cn[:, j]
    a, b = a.ravel(), b.ravel()
    ok = a != b
    if cn.min(initial=0) < 0:
        ok &= (a >= 0) & (b >= 0)
    if not ok.all():
        a, b = a[ok], b[ok]
    return a, b


class InterfaceProblem:
    """Assembled interface system of a decomposed lattice, resident on one GPU."""

    def __init__(self, ctx: L.Context, cell_nodes, n_interface_nodes, S):
        """cell_nodes: int [n_cells, n_bnd_nodes] interface node index of each local boundary node
        (``Point.index_boundary`` in ``cell.node_in_order_simulation`` order); S: device tensor
        [n_cells, nb, nb] or [nb, nb] (one Schur matrix shared by identical cells)."""
        import torch
        self.torch, self.ctx = torch, ctx
        dev = ctx.device
        self.n_nodes = int(n_interface_nodes)
        self.cell_nodes = torch.from_numpy(np.ascontiguousarray(cell_nodes, dtype=np.int32)).to(dev)
        cn = self.cell_nodes
        nc, nbn = int(cn.shape[0]), int(cn.shape[1])
        # pattern input: all unordered node pairs of every cell as virtual 2-node elements, generated on the device
        # (pair-major like cell_pair_elements; 6 M pairs at BASELINE config 3: 35 ms of host work + upload gone)
        ia, ib = np.triu_indices(nbn, k=1)
        a = cn[:, torch.from_numpy(ia).to(dev)].t().reshape(-1)
        b = cn[:, torch.from_numpy(ib).to(dev)].t().reshape(-1)
        ok = (a != b) & (a >= 0) & (b >= 0)
        if not bool(ok.all()):
            a, b = a[ok], b[ok]
        self.rowptr, self.colidx = ctx.bsr_pattern(a.contiguous(), b.contiguous(), self.n_nodes)
        del a, b, ok
        self._build_plan()
        self.S = S
        self.vals = self.assemble(S)

    def _build_plan(self):
        """Assembly plan of ``lat_assemble_cells_bsr_plan``: the contributions (cell, a, b) of every BSR block, sorted by
        block (cell-major inside a block, so the summation order is fixed).  Depends on the pattern only."""
        torch = self.torch
        cn, nn = self.cell_nodes.long(), self.n_nodes
        nc, nbn = int(cn.shape[0]), int(cn.shape[1])
        nnzb = int(self.colidx.numel())
        row = torch.repeat_interleave(torch.arange(nn, device=cn.device), (self.rowptr[1:] - self.rowptr[:-1]).long())
        bkey = row * nn + self.colidx.long()                       # ascending: rows ascending, columns sorted inside a row
        ckey = (cn[:, :, None] * nn + cn[:, None, :]).reshape(-1)  # flat index = (c * nbn + a) * nbn + b
        valid = ((cn[:, :, None] >= 0) & (cn[:, None, :] >= 0)).reshape(-1)
        d = torch.arange(nc * nbn * nbn, device=cn.device)
        if not bool(valid.all()):
            d, ckey = d[valid], ckey[valid]
        blk = torch.searchsorted(bkey, ckey)
        if bool((blk >= nnzb).any()) or not bool((bkey[blk.clamp_max(nnzb - 1)] == ckey).all()):
            raise L.LatticeB200Error("interface pattern does not contain every (node, node) pair of the cells")
        order = torch.argsort(blk, stable=True)
        self.plan_contrib = d[order].contiguous()
        ptr = torch.zeros(nnzb + 1, dtype=torch.int64, device=cn.device)
        ptr[1:] = torch.cumsum(torch.bincount(blk, minlength=nnzb), 0)
        if int(ptr[-1]) >= 2 ** 31:
            raise L.LatticeB200Error("assembly plan exceeds int32 offsets")
        self.plan_ptr = ptr.to(torch.int32).contiguous()

    def assemble(self, S, out=None):
        """K_G = sum_c P_c^T S_c P_c for new Schur matrices on the same pattern (a design iteration): plan-driven gather,
        no atomics, bit-reproducible.  S: [n_cells, nb, nb] or [nb, nb] (shared)."""
        self.S = S
        self.vals = self.ctx.assemble_cells_plan(S, int(self.cell_nodes.shape[1]), self.plan_ptr, self.plan_contrib, out=out)
        return self.vals

    def solve(self, fixed, g, f, tol=1e-10, maxiter=200000, precond=L.PC_BLOCK6, two_level=None, xyz=None):
        """``two_level`` (True or a number of aggregates) with ``xyz`` [n_interface_nodes, 3]: block-Jacobi + rigid-body-mode
        coarse space (coarse.TwoLevel) -- the role SuperLU's factorisation of this matrix plays in the reference
        (lattice_sim.py:1333-1415)."""
        import contextlib
        torch, ctx = self.torch, self.ctx
        dev = ctx.device
        t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
        fd, gd, fv = t(fixed, np.uint8), t(g, np.float64), t(f, np.float64)
        vbc, b = ctx.apply_dirichlet(self.rowptr, self.colidx, self.vals, fd, gd, fv)
        scope = contextlib.nullcontext()
        if two_level:
            from . import coarse
            if xyz is None:
                raise ValueError("two_level needs the interface node coordinates (xyz)")
            c = t(np.asarray(xyz, dtype=np.float64).T, np.float64)
            n_agg = coarse.default_aggregates(self.n_nodes) if two_level is True else int(two_level)
            scope = coarse.TwoLevel(ctx, c[0], c[1], c[2], fd, self.rowptr, self.colidx, vbc, n_agg)
        with scope:
            u, info = ctx.pcg(self.rowptr, self.colidx, vbc, b, tol=tol, maxiter=maxiter, precond=precond)
        ctx.set_dirichlet_values(fd, gd, u)
        R = ctx.spmv(self.rowptr, self.colidx, self.vals, u)
        return u, R, info, b

    def matvec_free(self, x_free, gidx, u_fixed=None):
        """The reference's interface operator on the FREE-DOF vector (lattice_sim.py:1180-1252)."""
        return self.ctx.ddm_matvec(self.S, gidx, x_free, u_fixed=u_fixed)


def interface_from_lattice(lattice, ctx=None):
    """Interface problem of a decomposed lattice from its (reference or duck-typed) object graph: needs
    ``cell.schur_complement`` on every cell and ``Point.index_boundary`` on the cell-boundary nodes.
    Returns (InterfaceProblem, {index_boundary: Point}, fixed, g, f) with 6 DOFs per interface node."""
    import torch
    ctx = ctx or L.default_context()
    cells = list(lattice.cells)
    for c in cells:
        if c.node_in_order_simulation is None:
            c.define_node_order_to_simulate()
    nbn = max(len(c.node_in_order_simulation) for c in cells)
    n_int = int(lattice.max_index_boundary) + 1
    cell_nodes = np.full((len(cells), nbn), -1, dtype=np.int32)
    # one device copy per UNIQUE Schur matrix would do; the per-cell stack keeps lat_assemble_cells_bsr's layout simple
    S = np.zeros((len(cells), 6 * nbn, 6 * nbn))
    pts = {}
    for k, c in enumerate(cells):
        nb = len(c.node_in_order_simulation)
        for a, p in enumerate(c.node_in_order_simulation):
            cell_nodes[k, a] = p.index_boundary
            pts[p.index_boundary] = p
        S[k, : 6 * nb, : 6 * nb] = np.asarray(c.schur_complement)
    fixed = np.zeros(6 * n_int, dtype=np.uint8)
    g = np.zeros(6 * n_int)
    f = np.zeros(6 * n_int)
    for ib, p in pts.items():
        for d in range(NDOF):
            if p.fixed_DOF[d]:
                fixed[6 * ib + d] = 1
                g[6 * ib + d] = p.displacement_vector[d]
            f[6 * ib + d] = float(p.applied_force[d])   # DDM applies every component once (lattice_sim.py:567-632)
    prob = InterfaceProblem(ctx, cell_nodes, n_int, torch.from_numpy(S).to(ctx.device))
    return prob, pts, fixed, g, f


def free_dof_map(lattice, pts, n_free=None):
    """Position (6 * index_boundary + d) of every free interface DOF in the reference's free-DOF numbering
    (``Point.global_free_DOF_index``, set by ``LatticeSim.set_global_free_DOF_index``, lattice_sim.py:654-669)."""
    pairs = []
    for ib, p in pts.items():
        for d in range(NDOF):
            if not p.fixed_DOF[d]:
                pairs.append((int(p.global_free_DOF_index[d]), 6 * ib + d))
    n_free = len(pairs) if n_free is None else int(n_free)
    out = np.full(n_free, -1, dtype=np.int64)
    for k, pos in pairs:
        out[k] = pos
    if (out < 0).any():
        raise ValueError("free-DOF numbering of the lattice is incomplete: call set_global_free_DOF_index() first")
    return out


TWO_LEVEL_AUTO_NODES = 20000     # interface systems from this size on get the coarse space by default


def solve_DDM_B200(lattice, tol=1e-10, maxiter=200000, ctx=None, two_level="auto"):
    """Drop-in for ``LatticeSim.solve_DDM()`` -> (xsol, info, global_displacement_index, b)
    (lattice_sim.py:1111-1176).

    Needs ``cell.schur_complement`` on every cell (``calculate_schur_complement_cells``, which calls the
    patched ``get_schur_complement``).  Leaves displacements and reactions on the boundary ``Point``s and, like
    the reference, the free-DOF numbering on the lattice (``define_free_DOF`` / ``set_global_free_DOF_index``
    when the object offers them)."""
    for name in ("define_free_DOF", "set_global_free_DOF_index"):
        fn = getattr(lattice, name, None)
        if callable(fn):
            fn()
    prob, pts, fixed, g, f = interface_from_lattice(lattice, ctx)
    if two_level == "auto":
        # the reference always preconditions this solve with a factorisation of the interface matrix (lattice_sim.py:
        # 1333-1415); the coarse space plays that role from the size on where its set-up pays (config 3: 797 -> 116 it)
        two_level = prob.n_nodes >= TWO_LEVEL_AUTO_NODES
    xyz = None
    if two_level:
        xyz = np.zeros((prob.n_nodes, 3))
        for ib, p in pts.items():
            xyz[ib] = (float(p.x), float(p.y), float(p.z))
    u, R, info, b = prob.solve(fixed, g, f, tol=tol, maxiter=maxiter, two_level=two_level, xyz=xyz)
    if info["info"] not in (0, 5):
        import warnings
        warnings.warn(f"solve_DDM_B200: interface PCG stopped with info={info['info']} (relres {info['relres']:.2e}, "
                      f"true {info['true_relres']:.2e})")
    uh, Rh = u.cpu().numpy().reshape(-1, 6), R.cpu().numpy().reshape(-1, 6)
    for ib, p in pts.items():
        p.displacement_vector[:] = [float(v) for v in uh[ib]]
        p.reaction_force_vector = [float(v) for v in Rh[ib]]
    xsol, idx = lattice.get_global_displacement()
    free = fixed.reshape(-1) == 0
    code = 0 if info["info"] in (0, 5) else int(info["info"])      # the reference's 0 / 1 / 2 convention
    return xsol, code, lattice.global_displacement_index, b.cpu().numpy()[free]


def compliance_gradient_cells(lattice, ctx=None, adjoint=None):
    """q[c, j] = u_c^T (dS_c/dr_j) u_c for every cell and geometry slot on the device (``lat_cell_quadform``) from the
    boundary displacements on the ``Point``s and ``cell.schur_complement_gradient`` -- the inner term of
    ``LatticeOpti.calculate_gradient`` (lattice_opti.py:752-761).  ``adjoint``: optional per-cell list of lambda_c."""
    import torch
    ctx = ctx or L.default_context()
    cells = list(lattice.cells)
    n_geom = max(len(getattr(c, "schur_complement_gradient", None) or []) for c in cells)
    if n_geom == 0:
        raise ValueError("no cell carries schur_complement_gradient: enable_gradient_computing was off")
    nb = max(6 * len(c.node_in_order_simulation) for c in cells)
    uniq, mats = {}, []
    index = np.full((len(cells), n_geom), -1, dtype=np.int32)
    U = np.zeros((len(cells), nb))
    V = None if adjoint is None else np.zeros((len(cells), nb))
    for k, c in enumerate(cells):
        if c.node_in_order_simulation is None:
            c.define_node_order_to_simulate()
        uc = np.asarray(c.get_displacement_at_nodes(c.node_in_order_simulation), dtype=np.float64).ravel()
        U[k, : uc.size] = uc
        if V is not None:
            V[k, : uc.size] = np.asarray(adjoint[k], dtype=np.float64).ravel()
        for j, dS in enumerate(getattr(c, "schur_complement_gradient", None) or []):
            key = id(dS)
            if key not in uniq:
                M = np.zeros((nb, nb))
                a = np.asarray(dS, dtype=np.float64)
                M[: a.shape[0], : a.shape[1]] = a
                uniq[key] = len(mats)
                mats.append(M)
            index[k, j] = uniq[key]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    q = ctx.cell_quadform(t(np.stack(mats)), t(index), t(U), None if V is None else t(V))
    return q.cpu().numpy()


def regular_bcc_interface(ctx, n_cells, cell_radii, elements_per_strut, young, nu, kappa=0.9):
    """Interface problem of a regular BCC lattice with one radius per cell (BASELINE configs[3]: BCC 60^3 with per-cell
    radii), built without an object graph: batched per-cell condensation (star-cell kernel) + the assembled interface
    operator over the (n+1)^3 cell corners.  Interface node (i, j, k) has index (i (ny+1) + j)(nz+1) + k.
    Returns (InterfaceProblem, corner_xyz [n_corners, 3], timings dict in ms)."""
    import torch
    from .mesh import synthetic_lattice
    from .schur import bcc_cell_order_nodes, synthetic_cell_batch
    nx, ny, nz = (int(v) for v in n_cells)
    radii = np.asarray(cell_radii, dtype=np.float64).ravel()
    if radii.shape[0] != nx * ny * nz:
        raise ValueError("one radius per cell, in cell-index (i-major) order")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e = [ev() for _ in range(4)]
    e[0].record()
    batch, _ = synthetic_cell_batch(ctx, "BCC", radii, elements_per_strut, young, nu)
    e[1].record()
    S = batch.schur()
    e[2].record()
    unit = synthetic_lattice("BCC", (1, 1, 1), [1.0])
    order = bcc_cell_order_nodes(unit.pxyz, (0, 1, 0, 1, 0, 1))
    off = np.rint(unit.pxyz[order]).astype(np.int64)                          # (8, 3) corner offsets in boundary order
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()                            # i-major == cell.index order
    cell_nodes = ((ci[:, None] + off[None, :, 0]) * (ny + 1) + (cj[:, None] + off[None, :, 1])) * (nz + 1) + (ck[:, None] + off[None, :, 2])
    prob = InterfaceProblem(ctx, cell_nodes.astype(np.int32), (nx + 1) * (ny + 1) * (nz + 1), S)
    e[3].record()
    torch.cuda.synchronize()
    gi, gj, gk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    corner_xyz = np.stack([gi.ravel(), gj.ravel(), gk.ravel()], axis=1).astype(np.float64)
    return prob, corner_xyz, dict(setup_ms=e[0].elapsed_time(e[1]), condense_ms=e[1].elapsed_time(e[2]),
                                  interface_assembly_ms=e[2].elapsed_time(e[3]))
