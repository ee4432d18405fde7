"""Multi-GPU sharding of the full-lattice solve: 1-D slab partition by x (cells are numbered
i-major in the reference, ``pyLatticeDesign/lattice.py:448-453``, so a slab of consecutive
x-layers is a contiguous cell-index range), one process per GPU.

Host side (numpy, this file): which nodes a rank owns, which neighbours' nodes it needs as
ghosts, local renumbering [owned | ghosts by owner rank, then global id], send lists.  Both
sides of every exchange derive their lists from the same global connectivity, so no
communication is needed to set the halos up.

Device side (``liblattice_b200.so``): complete block rows of the owned nodes are assembled
locally from the elements that touch them (a one-strut-deep overlap instead of exchanging
partial rows); PCG exchanges the ghost part of z with ncclSend/ncclRecv and all-reduces the dot
products (``lat_pcg_bsr_dist``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .mesh import NDOF, BeamMesh


@dataclass
class SlabPartition:
    rank: int
    world: int
    owned: np.ndarray        # global node ids owned by this rank (ascending)
    ghosts: np.ndarray       # global node ids of the ghosts, ordered by owner rank then global id
    ghost_owner: np.ndarray  # owner rank of each ghost
    local_elems: np.ndarray  # global element ids assembled on this rank (ascending)
    peers: list              # neighbour ranks (ascending)
    send_lists: list         # per peer: LOCAL ids (owned) to send, ordered by global id
    recv_counts: list        # per peer: number of ghosts owned by that peer
    g2l: dict | None = None

    @property
    def n_owned(self):
        return int(self.owned.shape[0])

    @property
    def n_local(self):
        return int(self.owned.shape[0] + self.ghosts.shape[0])

    @property
    def local_nodes(self):
        return np.concatenate([self.owned, self.ghosts])


def node_owner_by_x(x, bounds):
    """Owner rank of each node from its x coordinate: rank p owns bounds[p] <= x < bounds[p+1]
    (the last rank also owns x == bounds[-1])."""
    r = np.searchsorted(bounds, x, side="right") - 1
    return np.clip(r, 0, len(bounds) - 2).astype(np.int32)


def slab_bounds(x_min, x_max, world, n_layers=None, cell=1.0):
    """Slab boundaries on cell-layer planes: n_layers cells of size ``cell`` split as evenly as possible
    (explicit variant; :func:`slab_bounds_from_nodes` needs no cell size)."""
    if n_layers is None:
        n_layers = int(round((x_max - x_min) / cell))
    if world > n_layers:
        raise ValueError(f"slab partition: {world} ranks for {n_layers} cell layers")
    cuts = [x_min + cell * ((n_layers * p) // world) for p in range(world)] + [x_max]
    return np.array(cuts, dtype=np.float64)


def slab_bounds_from_nodes(x, world):
    """Slab boundaries derived from the data: cuts fall on node x-planes such that every rank owns about the same
    number of nodes.  Independent of the cell size and of where the lattice starts; never produces an empty slab
    as long as there are at least ``world`` distinct x-planes (ValueError otherwise, identically on every rank
    because every rank evaluates the same global array)."""
    ux, cnt = np.unique(np.round(np.asarray(x, dtype=np.float64), 9), return_counts=True)
    if ux.size < world:
        raise ValueError(f"slab partition: {world} ranks but only {ux.size} distinct x-planes of nodes")
    cum = np.cumsum(cnt)
    total = int(cum[-1])
    first = [0]                                   # index of the first plane of every rank
    for p in range(1, world):
        k = int(np.searchsorted(cum, p * total / world, side="left")) + 1   # planes [0, k) hold >= p/world of the nodes
        k = max(k, first[-1] + 1)                 # at least one plane per rank ...
        k = min(k, ux.size - (world - p))         # ... and one left for each rank behind
        first.append(k)
    # a cut half-way towards the previous plane keeps rounding noise in x away from the comparison
    cuts = [ux[0]] + [0.5 * (ux[k - 1] + ux[k]) for k in first[1:]] + [ux[-1]]
    return np.array(cuts, dtype=np.float64)


def partition_slab(mesh: BeamMesh, rank: int, world: int, bounds=None, owner=None, check=True) -> SlabPartition:
    if owner is None:
        if bounds is None:
            bounds = slab_bounds_from_nodes(mesh.x, world)
        owner = node_owner_by_x(mesh.x, bounds)
    n_per_rank = np.bincount(owner, minlength=world)
    if check and (n_per_rank[:world] == 0).any():           # same verdict on every rank -> nobody enters a collective
        raise ValueError(f"slab partition: ranks {np.flatnonzero(n_per_rank[:world] == 0).tolist()} own no node "
                         f"(bounds {None if bounds is None else list(bounds)})")
    o0, o1 = owner[mesh.en0], owner[mesh.en1]
    mine = (o0 == rank) | (o1 == rank)
    local_elems = np.flatnonzero(mine)
    owned = np.flatnonzero(owner == rank)
    ends = np.concatenate([mesh.en0[local_elems], mesh.en1[local_elems]])
    ghosts = np.unique(ends[owner[ends] != rank])
    gown = owner[ghosts]
    order = np.lexsort((ghosts, gown))
    ghosts, gown = ghosts[order], gown[order]
    peers = sorted(set(gown.tolist()))
    # what I must send to peer q: my owned nodes that are element-neighbours of a node owned by q
    e_cross0 = (o0 == rank) & (o1 != rank)
    e_cross1 = (o1 == rank) & (o0 != rank)
    send_pairs = np.concatenate([np.stack([mesh.en0[e_cross0], o1[e_cross0]], 1),
                                 np.stack([mesh.en1[e_cross1], o0[e_cross1]], 1)], axis=0)
    peers = sorted(set(peers) | set(send_pairs[:, 1].tolist()))
    g2l = np.full(mesh.n_nodes, -1, dtype=np.int64)          # global -> local id of the owned nodes (no per-node dict)
    g2l[owned] = np.arange(owned.shape[0])
    send_lists, recv_counts = [], []
    for q in peers:
        nodes = np.unique(send_pairs[send_pairs[:, 1] == q, 0])
        send_lists.append(g2l[nodes].astype(np.int32))
        recv_counts.append(int((gown == q).sum()))
    # every rank can evaluate every other rank's neighbour count from the same arrays: fail everywhere, not on one rank
    cross = o0 != o1
    pair = np.unique(np.stack([np.minimum(o0[cross], o1[cross]), np.maximum(o0[cross], o1[cross])], 1), axis=0)
    n_nb = np.bincount(pair.ravel(), minlength=world) if pair.size else np.zeros(world, dtype=np.int64)
    if check and n_nb.max(initial=0) > 4:
        raise ValueError(f"slab partition: a rank would have {int(n_nb.max())} neighbours (slabs thinner than one strut); "
                         "use fewer ranks")
    return SlabPartition(rank, world, owned, ghosts, gown, local_elems, peers, send_lists, recv_counts)


def local_mesh(mesh: BeamMesh, part: SlabPartition) -> BeamMesh:
    """Sub-mesh of a rank in local numbering [owned | ghosts]."""
    nodes = part.local_nodes
    g2l = np.full(mesh.n_nodes, -1, dtype=np.int64)
    g2l[nodes] = np.arange(nodes.shape[0])
    e = part.local_elems
    return BeamMesh(x=mesh.x[nodes].copy(), y=mesh.y[nodes].copy(), z=mesh.z[nodes].copy(),
                    en0=g2l[mesh.en0[e]].astype(np.int32), en1=g2l[mesh.en1[e]].astype(np.int32),
                    rad=mesh.rad[e].copy(), beam_of_elem=mesh.beam_of_elem[e].copy(), chain=mesh.chain[e].copy(),
                    n_points=int((nodes < mesh.n_points).sum()), point_index=nodes[nodes < mesh.n_points],
                    cell_of_elem=None if mesh.cell_of_elem is None else mesh.cell_of_elem[e].copy(),
                    meta={"global_nodes": nodes, "is_point": nodes < mesh.n_points})


def compression_bc_local(lm: BeamMesh, value=-0.01):
    """``mesh.compression_bc`` for a LOCAL slab mesh (lattice points are flagged in ``meta['is_point']``, not the
    first nodes): clamp Zmin, impose u_z = value on Zmax.  Valid for x-slabs: every slab sees both z extremes."""
    is_pt = lm.meta["is_point"]
    zp = lm.z[is_pt]
    zmin, zmax = zp.min(), zp.max()
    n = lm.n_dof
    fixed, g, f = np.zeros(n, dtype=np.uint8), np.zeros(n), np.zeros(n)
    bot = np.flatnonzero(is_pt & (lm.z == zmin))
    top = np.flatnonzero(is_pt & (lm.z == zmax))
    fixed[(bot[:, None] * NDOF + np.arange(NDOF)[None, :]).ravel()] = 1
    fixed[top * NDOF + 2] = 1
    g[top * NDOF + 2] = value
    return fixed, g, f


def local_dofs(part: SlabPartition):
    nodes = part.local_nodes
    return (nodes[:, None] * NDOF + np.arange(NDOF)[None, :]).ravel()


def slab_layers(n_layers, world):
    """Cell layers [i0, i1) of every rank: n_layers split as evenly as possible (ValueError when world > n_layers)."""
    if world > n_layers:
        raise ValueError(f"slab partition: {world} ranks for {n_layers} cell layers")
    return [((n_layers * p) // world, (n_layers * (p + 1)) // world) for p in range(world)]


def generate_slab(geom_types, n_cells, radii, elements_per_strut, rank, world, cell_size=(1.0, 1.0, 1.0),
                  grad_radius=None, cell_radii=None, device=None, _general=False):
    """Per-slab lattice generation: the rank builds ONLY its own cell layers plus one overlap layer on each side
    (host time and memory ~ 1/world of the full lattice) and returns (local BeamMesh in [owned | ghosts] numbering,
    SlabPartition).  No global numbering exists on this path: both sides of an exchange order the shared nodes by
    their rank in the generator's numbering, which is monotone in the reference's global numbering
    (lattice points by (x, y, z), strut-interior nodes beam-major), so send and receive lists agree without
    communication -- the same contract as :func:`partition_slab`.  ``device``: torch device for the numbering of the
    lattice points (``mesh._grid_lattice_torch``; bit-identical to the host path)."""
    from .mesh import synthetic_lattice, mesh_from_synthetic
    nx = int(n_cells[0])
    i0, i1 = slab_layers(nx, world)[rank]
    lat = synthetic_lattice(geom_types, n_cells, radii, cell_size=cell_size, grad_radius=grad_radius,
                            cell_radii=cell_radii, i_range=(i0 - 1, i1 + 1), device=device)
    mesh = mesh_from_synthetic(lat, elements_per_strut)
    if world == 1 and not _general:
        # one rank owns everything: the partition and the local numbering are the identity (no gathers, no copies)
        nodes = np.arange(mesh.n_nodes, dtype=np.int64)
        empty = np.zeros(0, dtype=np.int64)
        part = SlabPartition(0, 1, nodes, empty, empty, np.arange(mesh.n_elems, dtype=np.int64), [], [], [])
        lm = BeamMesh(x=mesh.x, y=mesh.y, z=mesh.z, en0=mesh.en0.astype(np.int32, copy=False), en1=mesh.en1.astype(np.int32, copy=False),
                      rad=mesh.rad, beam_of_elem=mesh.beam_of_elem, chain=mesh.chain, n_points=mesh.n_points,
                      point_index=nodes[: mesh.n_points], cell_of_elem=mesh.cell_of_elem,
                      meta={"global_nodes": nodes, "is_point": nodes < mesh.n_points})
        lm.meta["lattice"] = lat
        lm.meta["layers"] = (i0, i1)
        return lm, part
    cs = float(cell_size[0])
    # cell-plane coordinates exactly as the generator accumulates them (lattice.py:433-442)
    xs = np.concatenate([[0.0], np.cumsum(np.full(max(nx - 1, 0), cs))])[:nx]
    cuts = [xs[a] for a, _ in slab_layers(nx, world)] + [xs[nx - 1] + cs]
    bounds = np.array(cuts, dtype=np.float64)
    bounds[1:-1] -= 1e-9 * cs                      # a node ON a cut plane belongs to the upper rank on both sides
    owner = node_owner_by_x(mesh.x, bounds)
    part = partition_slab(mesh, rank, world, owner=owner, check=False)
    if any(abs(q - rank) != 1 for q in part.peers):
        raise ValueError("generate_slab: a strut spans more than one slab; use fewer ranks")
    lm = local_mesh(mesh, part)
    lm.meta["lattice"] = lat
    lm.meta["layers"] = (i0, i1)
    return lm, part


class DistributedFEM:
    """Slab-sharded assemble + solve.  Two constructors: the plain one takes the GLOBAL mesh (every rank passes the
    same arrays and keeps only its slab on the GPU; global DOF numbering available for gather / parity checks);
    :meth:`from_generator` builds only the rank's own slab (``generate_slab``)."""

    def __init__(self, ctx, mesh: BeamMesh, young, nu, rank, world, kappa=0.9, bounds=None, part=None, lmesh=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        if part is None:
            part = partition_slab(mesh, rank, world, bounds)
            lmesh = local_mesh(mesh, part)
            self.dofs = local_dofs(part)
            self.n_dof_global = mesh.n_dof
            self.n_elem_global = mesh.n_elems
        else:
            self.dofs = None
            self.n_dof_global = self.n_elem_global = None
        self.part, self.lmesh = part, lmesh
        self.young, self.nu, self.kappa = young, nu, kappa
        self.upload()

    @classmethod
    def from_generator(cls, ctx, geom_types, n_cells, radii, elements_per_strut, young, nu, rank, world, kappa=0.9,
                       cell_size=(1.0, 1.0, 1.0), grad_radius=None, cell_radii=None):
        import torch
        import torch.distributed as dist
        lm, part = generate_slab(geom_types, n_cells, radii, elements_per_strut, rank, world, cell_size, grad_radius,
                                 cell_radii, device=ctx.device)
        self = cls(ctx, None, young, nu, rank, world, kappa, part=part, lmesh=lm)
        # global sizes: owned nodes, and elements counted once (by the owner of their first node)
        cnt = torch.tensor([part.n_owned, int((lm.en0 < part.n_owned).sum())], dtype=torch.int64)
        if world > 1:
            cnt = cnt.to(ctx.device) if dist.get_backend() == "nccl" else cnt
            dist.all_reduce(cnt)
        self.n_dof_global, self.n_elem_global = 6 * int(cnt[0]), int(cnt[1])
        return self

    def upload(self):
        """Host -> device copy of the local mesh, halo lists and the BSR pattern of the local mesh."""
        import torch
        from . import lib as L
        self.torch, self.L = torch, L
        ctx = self.ctx
        dev = ctx.device
        t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
        m = self.lmesh
        self.x, self.y, self.z = t(m.x, np.float64), t(m.y, np.float64), t(m.z, np.float64)
        self.en0, self.en1, self.rad = t(m.en0, np.int32), t(m.en1, np.int32), t(m.rad, np.float64)
        self.n_owned, self.n_local = self.part.n_owned, self.part.n_local
        send = np.concatenate(self.part.send_lists) if self.part.send_lists else np.zeros(0, np.int32)
        self.send_idx = t(send, np.int32)
        self.halo = ctx.make_halo(self.part.peers, [len(s) for s in self.part.send_lists], self.part.recv_counts,
                                  self.send_idx, self.n_owned, self.n_local)
        # pattern over the local mesh; only the first n_owned block rows are complete and used
        self.rowptr, self.colidx = ctx.bsr_pattern(self.en0, self.en1, self.n_local)
        self.nnzb = int(self.colidx.numel())
        self.nnzb_owned = int(self.rowptr[self.n_owned].item())
        self.vals = None

    def set_bc_local(self, fixed, g, f):
        """Boundary conditions given directly in the LOCAL numbering [owned | ghosts] (the per-slab path)."""
        t = lambda a, d: self.torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(self.ctx.device)
        self.fixed_d, self.g_d, self.f_d = t(fixed, np.uint8), t(g, np.float64), t(f, np.float64)

    def set_bc(self, fixed, g, f):
        if self.dofs is None:
            raise ValueError("this DistributedFEM has no global numbering (from_generator): use set_bc_local")
        self.set_bc_local(fixed[self.dofs], g[self.dofs], f[self.dofs])

    def assemble(self, out=None):
        self.vals = self.ctx.assemble_bsr(self.x, self.y, self.z, self.en0, self.en1, self.rad, self.n_local, self.nnzb,
                                          self.young, self.nu, self.kappa, out=out)
        return self.vals

    def enable_p2p(self):
        """Switch the iteration's halo exchange / all-reduce from NCCL to NVLink peer memory written by our
        own kernels.  Every rank tells its neighbours where their data lands: the first local node index of
        the ghost segment owned by each peer."""
        import torch.distributed as dist
        torch = self.torch
        # my ghost segments: peer q's data starts at n_owned + sum(recv_counts of earlier peers)
        off, seg0 = self.n_owned, {}
        for q, rc in zip(self.part.peers, self.part.recv_counts):
            seg0[q] = off
            off += rc
        table = torch.full((self.world,), -1, dtype=torch.int64, device=self.ctx.device)
        for q, v in seg0.items():
            table[q] = v
        allt = [torch.zeros_like(table) for _ in range(self.world)]
        dist.all_gather(allt, table)          # allt[q][me] = where MY data lands on rank q
        dst0 = [int(allt[q][self.rank]) for q in self.part.peers]
        self.ctx.p2p_setup(self.n_local, self.part.peers, dst0)
        self.p2p = True

    def two_level(self, n_aggregates=None):
        """Rigid-body-mode coarse space of the sharded system (coarse.TwoLevel): boxes over the GLOBAL bounding box, so
        every rank numbers the aggregates alike; each rank forms the Galerkin product of its owned rows (the ghost
        columns included), the coarse matrix is summed over the ranks and every rank inverts it.  Needs
        ``set_bc_local`` first; a matrix-free user pays one temporary assembly."""
        import torch.distributed as dist
        from . import coarse
        torch, ctx = self.torch, self.ctx
        dev = ctx.device
        lo = torch.stack([self.x.min(), self.y.min(), self.z.min()])
        hi = torch.stack([self.x.max(), self.y.max(), self.z.max()])
        n_glob = torch.tensor([float(self.n_owned)], dtype=torch.float64, device=dev)
        if self.world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(n_glob)
        target = int(n_aggregates or coarse.default_aggregates(int(n_glob.item())))
        lo_h = lo.cpu().numpy()
        nb, ext = coarse.box_grid(lo_h, hi.cpu().numpy(), min(target, coarse.MAX_AGGREGATES))
        agg = coarse.box_index(self.x, self.y, self.z, lo_h, ext, nb)       # global box index, empty boxes kept
        n_agg = int(nb.prod())
        cen = coarse.box_centers(lo_h, ext, nb)
        vals = self.vals
        if vals is None:            # temporary assembly (the subclass' operator for the joint-only system)
            vals = self.assemble()
            self.vals = None
        allred = (lambda E: dist.all_reduce(E)) if self.world > 1 else None
        return coarse.TwoLevel(ctx, self.x, self.y, self.z, self.fixed_d, self.rowptr, self.colidx, vals, agg=agg, n_agg=n_agg,
                               n_owned=self.n_owned, centers=torch.from_numpy(cen).to(dev), allreduce=allred)

    def _two_level_scope(self, two_level):
        import contextlib
        if two_level is None or two_level is False:
            return contextlib.nullcontext()
        from . import coarse
        if isinstance(two_level, coarse.TwoLevel):
            return two_level
        return self.two_level(None if two_level is True else int(two_level))

    def solve(self, tol=1e-8, maxiter=200000, precond=2, vals_bc=None, b=None, u=None, check_every=0, profile_iters=0,
              overlap=False, fused_halo=True, persistent=True, two_level=None):
        """Returns (u_local [6 n_local] incl. ghosts, reactions on owned rows, info)."""
        torch, ctx = self.torch, self.ctx
        if self.vals is None:
            self.assemble()
        if vals_bc is None:
            vals_bc = torch.empty_like(self.vals)
        if b is None:
            b = torch.empty(6 * self.n_local, dtype=torch.float64, device=ctx.device)
        if u is None:
            u = torch.empty(6 * self.n_local, dtype=torch.float64, device=ctx.device)
        L = self.L
        ctx.check(ctx.lib.lat_apply_dirichlet(ctx.h, L._ptr(self.rowptr), L._ptr(self.colidx), self.n_local,
                                              L._ptr(self.vals), L._ptr(self.fixed_d), L._ptr(self.g_d),
                                              L._ptr(self.f_d), L._ptr(vals_bc), L._ptr(b)))
        with self._two_level_scope(two_level):
            u, info = ctx.pcg_dist(self.rowptr, self.colidx, vals_bc, self.halo, b, u, tol=tol, maxiter=maxiter,
                                   precond=precond, check_every=check_every, p2p=getattr(self, "p2p", False),
                                   profile_iters=profile_iters, overlap=overlap, fused_halo=fused_halo, persistent=persistent)
        ctx.set_dirichlet_values(self.fixed_d, self.g_d, u)
        ctx.halo_exchange(self.halo, u)
        R = ctx.spmv(self.rowptr, self.colidx, self.vals, u)     # rows >= n_owned are partial: ignore
        return u, R, info

    def solve_matrix_free(self, tol=1e-8, maxiter=200000, precond=2, b=None, u=None, check_every=0, profile_iters=0,
                          want_reactions=True, overlap=False, fused_halo=True, two_level=None):
        """:meth:`solve` without an assembled matrix (csrc/matfree.cuh): the operator is regenerated from the
        local mesh in every product; halo exchange and all-reduce are unchanged."""
        torch, ctx = self.torch, self.ctx
        if b is None:
            b = torch.empty(6 * self.n_local, dtype=torch.float64, device=ctx.device)
        if u is None:
            u = torch.empty(6 * self.n_local, dtype=torch.float64, device=ctx.device)
        ctx.matfree_setup(self.x, self.y, self.z, self.en0, self.en1, self.rad, self.n_local, self.young, self.nu,
                          self.kappa, fixed=self.fixed_d)
        ctx.matfree_rhs(self.g_d, self.f_d, out=b)        # rows >= n_owned are partial: never read
        with self._two_level_scope(two_level):
            u, info = ctx.pcg_matfree_dist(self.halo, b, u, tol=tol, maxiter=maxiter, precond=precond,
                                           check_every=check_every, p2p=getattr(self, "p2p", False),
                                           profile_iters=profile_iters, overlap=overlap, fused_halo=fused_halo)
        ctx.set_dirichlet_values(self.fixed_d, self.g_d, u)
        ctx.halo_exchange(self.halo, u)
        R = ctx.matfree_apply(u, eliminated=False) if want_reactions else None
        return u, R, info

    def compliance_gradient(self, u_local, group_global, n_groups, chain_global=None):
        """g[p] = -sum_e chain_e u_e^T dK_e/dr u_e over ALL elements of the lattice (lattice_opti.py:746-841).
        Every element is counted on exactly one rank (the owner of its first node); the per-rank partial
        gradients are summed with one all-reduce.  ``u_local`` must carry valid ghosts (``solve`` returns it so)."""
        torch = self.torch
        e = self.part.local_elems
        owner_first = self.lmesh.en0 < self.n_owned            # first node owned by this rank
        grp = np.where(owner_first, np.asarray(group_global)[e], -1).astype(np.int32)
        dev = self.ctx.device
        ch = None
        if chain_global is not None:
            ch = torch.from_numpy(np.ascontiguousarray(np.asarray(chain_global)[e], dtype=np.float64)).to(dev)
        g = self.ctx.compliance_grad(self.x, self.y, self.z, self.en0, self.en1, self.rad, torch.from_numpy(grp).to(dev),
                                     n_groups, u_local, self.young, self.nu, self.kappa, chain=ch)
        return self.ctx.allreduce_sum(g)

    def gather_owned(self, vec_local):
        """All-gather the owned part of a local vector into a global numpy vector (test / write-back helper)."""
        import torch.distributed as dist
        torch = self.torch
        own = vec_local[: 6 * self.n_owned].contiguous()
        counts = torch.tensor([own.numel()], device=own.device)
        allc = [torch.zeros_like(counts) for _ in range(self.world)]
        dist.all_gather(allc, counts)
        mx = int(max(int(c) for c in allc))
        pad = torch.zeros(mx, dtype=own.dtype, device=own.device)
        pad[: own.numel()] = own
        bufs = [torch.zeros_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad)
        ids = torch.from_numpy(self.part.owned).to(own.device)
        idpad = torch.full((mx // 6,), -1, dtype=torch.int64, device=own.device)
        idpad[: ids.numel()] = ids
        idb = [torch.zeros_like(idpad) for _ in range(self.world)]
        dist.all_gather(idb, idpad)
        out = np.zeros(self.n_dof_global)
        for q in range(self.world):
            k = int(allc[q]) // 6
            gi = idb[q][:k].cpu().numpy()
            out.reshape(-1, 6)[gi] = bufs[q][: 6 * k].cpu().numpy().reshape(-1, 6)
        return out


# ---------------------------------------------------------------------------------------------------------------
# joint-only (strut-condensed) system, sharded
# ---------------------------------------------------------------------------------------------------------------
def strut_topology(mesh: BeamMesh):
    """Struts of a subdivided mesh in mesh.py numbering (lattice points first, elements beam-major from point1 to
    point2): element range ``ptr[s]:ptr[s+1]`` and the two end joints ``a[s]``, ``b[s]`` of every strut."""
    starts = np.flatnonzero(mesh.en0 < mesh.n_points)
    ptr = np.r_[starts, mesh.n_elems]
    a, b = mesh.en0[starts].astype(np.int64), mesh.en1[ptr[1:] - 1].astype(np.int64)
    if (b >= mesh.n_points).any() or (np.diff(ptr) < 1).any():
        raise ValueError("mesh is not beam-major between lattice points")
    return ptr.astype(np.int64), a, b


def joint_mesh(mesh: BeamMesh, sa, sb) -> BeamMesh:
    """The lattice points as nodes and the struts as elements: what the joint-only system is partitioned on."""
    npnt = mesh.n_points
    ns = sa.shape[0]
    return BeamMesh(x=mesh.x[:npnt].copy(), y=mesh.y[:npnt].copy(), z=mesh.z[:npnt].copy(), en0=sa.astype(np.int32),
                    en1=sb.astype(np.int32), rad=np.zeros(ns), beam_of_elem=np.arange(ns), chain=np.ones(ns), n_points=npnt,
                    point_index=np.arange(npnt), cell_of_elem=None, meta={})


def local_strut_mesh(mesh: BeamMesh, ptr, part: SlabPartition):
    """Sub-mesh a rank needs for the joint-only system: its local struts (``part.local_elems`` of the joint mesh) with
    all their elements.  Local node numbering: [owned joints | ghost joints | strut-interior nodes, strut-major].
    Returns dict(xyz, len0, len1, rad, chain_ptr, sa, sb, max_len, interior_global)."""
    struts = part.local_elems
    joints = part.local_nodes
    g2l = np.full(mesh.n_nodes, -1, dtype=np.int64)
    g2l[joints] = np.arange(joints.shape[0])
    cnt = (ptr[struts + 1] - ptr[struts]).astype(np.int64)
    cptr = np.r_[0, np.cumsum(cnt)]
    # global element ids of the local struts, strut-major
    elems = (np.repeat(ptr[struts] - cptr[:-1], cnt) + np.arange(cptr[-1])).astype(np.int64)
    e0, e1 = mesh.en0[elems].astype(np.int64), mesh.en1[elems].astype(np.int64)
    interior = e1[e1 >= mesh.n_points]                    # every interior node is the second node of exactly one element
    g2l[interior] = joints.shape[0] + np.arange(interior.shape[0])
    nodes = np.concatenate([joints, interior])
    return dict(xyz=np.stack([mesh.x[nodes], mesh.y[nodes], mesh.z[nodes]], axis=1), len0=g2l[e0].astype(np.int32),
                len1=g2l[e1].astype(np.int32), rad=mesh.rad[elems].copy(), chain_ptr=cptr.astype(np.int32),
                max_len=int(cnt.max(initial=1)), interior_global=interior, elems_global=elems)


class DistributedJointFEM(DistributedFEM):
    """:class:`DistributedFEM` on the exact joint-only system (every strut condensed onto its two lattice points,
    ``lat_assemble_bsr_struts``): the slab partition, halo lists, peer-memory exchange and PCG are those of the base
    class applied to the JOINT mesh; each rank condenses the struts that touch one of its joints.  Loads and
    constraints live on lattice points (what the reference applies, full_scale_lattice_simulation.py:77-153).
    ``solve`` returns the joint displacements / reactions in the local joint numbering [owned | ghosts];
    :meth:`recover_full_field` back-substitutes the interior nodes of the rank's struts."""

    def __init__(self, ctx, mesh: BeamMesh, young, nu, rank, world, kappa=0.9, bounds=None):
        ptr, sa, sb = strut_topology(mesh)
        jm = joint_mesh(mesh, sa, sb)
        part = partition_slab(jm, rank, world, bounds)
        self.full = local_strut_mesh(mesh, ptr, part)
        self.n_dof_full_global = mesh.n_dof
        super().__init__(ctx, None, young, nu, rank, world, kappa, part=part, lmesh=local_mesh(jm, part))
        self.dofs = local_dofs(part)                       # joint DOFs = the first 6 n_points DOFs of the full mesh
        self.n_dof_global, self.n_elem_global = 6 * mesh.n_points, mesh.n_elems
        t = lambda a, d: self.torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)
        f = self.full
        ne = f["len0"].shape[0]
        self.f_xyz, self.f_len0, self.f_len1, self.f_rad = t(f["xyz"], np.float64), t(f["len0"], np.int32), t(f["len1"], np.int32), t(f["rad"], np.float64)
        self.f_ptr, self.f_elem, self.f_flip = t(f["chain_ptr"], np.int32), t(np.arange(ne), np.int32), t(np.zeros(ne), np.int32)

    def assemble(self, out=None):
        self.vals = self.ctx.assemble_bsr_struts(self.f_xyz, self.f_len0, self.f_len1, self.f_rad, self.f_ptr, self.f_elem,
                                                 self.f_flip, self.n_local, self.nnzb, self.young, self.nu, self.kappa, out=out)
        return self.vals

    def solve_matrix_free(self, *a, **k):
        raise NotImplementedError("the joint-only system is an assembled operator")

    def compliance_gradient(self, *a, **k):
        raise NotImplementedError("use recover_full_field + the element-form gradient of the full mesh")

    def recover_full_field(self, u_joints_local):
        """Displacements of ALL local nodes ([owned joints | ghost joints | interior nodes of the local struts]) from
        the joint solution (ghost joints valid, as ``solve`` returns them): ``lat_strut_recover``."""
        torch, ctx = self.torch, self.ctx
        nloc = int(self.f_xyz.shape[0])
        u_full = torch.zeros(6 * nloc, dtype=torch.float64, device=ctx.device)
        u_full[: 6 * self.n_local] = u_joints_local
        if nloc > self.n_local:
            ctx.strut_recover(self.f_xyz, self.f_len0, self.f_len1, self.f_rad, self.f_ptr, self.f_elem, self.f_flip,
                              self.en0, self.en1, self.full["max_len"], self.young, self.nu, self.kappa, u_joints_local, u_full)
        return u_full
