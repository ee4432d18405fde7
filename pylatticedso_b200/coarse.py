"""Two-level preconditioner: host policy around csrc/coarse.cuh (aggregates, coarse factorisation, activation).

The reference preconditions its interface PCG with SuperLU's factorisation of the whole interface matrix
(``LatticeSim.define_preconditioner`` / ``build_preconditioner``, lattice_sim.py:1333-1415): a host factorisation that
cannot run inside a device-resident iteration.  Here the 6x6 block-Jacobi preconditioner is complemented by a coarse
space of rigid-body modes per aggregate of nodes (boxes of the bounding box): ``M^-1 = D^-1 + Z E^+ Z^T``.

Device work: aggregate tables, ``E = Z^T A Z`` and the two per-iteration kernels are CUDA (``lat_coarse_*``).  The
small dense ``E`` (6 n_agg, a few thousand at most) is inverted once on the GPU through ``torch.linalg`` (cuSOLVER:
a plain library factorisation at set-up, like the reference's SuperLU call; not in the iteration).
"""
from __future__ import annotations

import numpy as np

from . import lib as L

MAX_AGGREGATES = 2048            # 6 * 2048 = 12288 coarse DOF: 1.2 GB of dense inverse


def default_aggregates(n_nodes):
    """About 500 nodes per aggregate, at most 1000 aggregates: beyond that the dense coarse factorisation (O(n_c^3), redone
    whenever the radii change) costs more than the iterations it saves (profiles/r02_two_level_ab.txt)."""
    return int(max(8, min(1000, round(n_nodes / 500.0))))


def box_grid(lo, hi, target):
    """Boxes per axis (numpy int64 [3]) of a regular grid over [lo, hi] with about ``target`` boxes, as cubic as the
    extents allow; flat axes get one box.  Pure host arithmetic: the ranks of a sharded solve call it with the GLOBAL
    bounds and get the same grid."""
    ext = np.maximum(np.asarray(hi, dtype=np.float64) - np.asarray(lo, dtype=np.float64), 1e-300)
    live = ext > 1e-9 * ext.max()
    h = (np.prod(ext[live]) / max(1, int(target))) ** (1.0 / max(1, int(live.sum())))
    return np.where(live, np.maximum(1, np.round(ext / h)), 1).astype(np.int64), ext


def box_index(x, y, z, lo, ext, nb):
    """Flat box index (torch int64) of every node: (i * nb_y + j) * nb_z + k, nodes on the upper faces in the last box."""
    ijk = []
    for k, c in enumerate((x, y, z)):
        ijk.append(((c - float(lo[k])) / float(ext[k]) * float(nb[k])).floor().long().clamp_(0, int(nb[k]) - 1))
    return (ijk[0] * int(nb[1]) + ijk[1]) * int(nb[2]) + ijk[2]


def box_centers(lo, ext, nb):
    """Centres of all boxes in flat-index order (numpy [prod(nb), 3])."""
    gi, gj, gk = np.meshgrid(np.arange(nb[0]), np.arange(nb[1]), np.arange(nb[2]), indexing="ij")
    return np.asarray(lo, dtype=np.float64)[None, :] + (np.stack([gi.ravel(), gj.ravel(), gk.ravel()], 1) + 0.5) * (ext / nb)[None, :]


def box_aggregates(x, y, z, target):
    """Aggregate index per node (torch int64 on the nodes' device) and the number of aggregates: a regular grid of
    boxes over the bounding box, box counts per axis proportional to the extents, empty boxes dropped."""
    import torch
    target = int(max(1, min(target, MAX_AGGREGATES)))
    lo = torch.stack([x.min(), y.min(), z.min()]).cpu().numpy()
    hi = torch.stack([x.max(), y.max(), z.max()]).cpu().numpy()
    nb, ext = box_grid(lo, hi, target)
    uniq, inv = torch.unique(box_index(x, y, z, lo, ext, nb), return_inverse=True)
    return inv, int(uniq.numel())


def invert_coarse(E):
    """Symmetric (pseudo-)inverse of the coarse matrix on the device.  Aggregates whose DOFs are all constrained give
    exactly zero rows / columns (their restricted residual is exactly zero as well): they get a unit diagonal.
    Cholesky when E is comfortably positive definite, otherwise an eigen-decomposition with the null space cut."""
    import torch
    E = 0.5 * (E + E.T)
    d = E.diagonal()
    dead = d <= 0.0
    if bool(dead.any()):
        E = E.clone()
        E[dead, :] = 0.0
        E[:, dead] = 0.0
        E.diagonal()[dead] = 1.0
    Lc, info = torch.linalg.cholesky_ex(E)
    ok = int(info) == 0
    if ok:
        dl = Lc.diagonal()
        ok = bool((dl.min() / dl.max()) ** 2 > 1e-11)
    if ok:
        Einv = torch.cholesky_inverse(Lc)
    else:
        w, V = torch.linalg.eigh(E)
        keep = w > 1e-11 * w.max()
        Vk = V[:, keep]
        Einv = (Vk / w[keep]) @ Vk.T
    Einv = 0.5 * (Einv + Einv.T)
    if bool(dead.any()):
        Einv[dead, :] = 0.0
        Einv[:, dead] = 0.0
    return Einv.contiguous()


class TwoLevel:
    """Coarse space of one system, resident in ``ctx``.  Use as a context manager around the solves it preconditions::

        with TwoLevel(ctx, x, y, z, fixed, rowptr, colidx, vals):
            u, info = ctx.pcg(rowptr, colidx, vbc, rhs, precond=L.PC_BLOCK6)     # info['two_level'] is True

    ``rowptr/colidx/vals``: the BSR matrix the coarse operator is the Galerkin projection of (eliminated or not: the
    constrained DOFs are masked out of the coarse space)."""

    def __init__(self, ctx: L.Context, x, y, z, fixed, rowptr, colidx, vals, n_aggregates=None, agg=None, n_agg=None,
                 n_owned=None, centers=None, allreduce=None):
        """Sharded systems (distributed.DistributedFEM.two_level): ``agg`` / ``n_agg`` = the GLOBAL aggregate of every
        local node, ``n_owned`` = the leading nodes this rank owns (the rest are ghosts), ``centers`` [n_agg, 3] common
        reference points, ``allreduce(E)`` sums the rank contributions of the coarse matrix in place."""
        import torch
        self.ctx = ctx
        if agg is None:
            agg, n_agg = box_aggregates(x, y, z, n_aggregates or default_aggregates(int(x.numel())))
        elif n_agg is None:
            n_agg = int(agg.max()) + 1
        n_nodes = int(x.numel())
        n_owned = n_nodes if n_owned is None else int(n_owned)
        own = agg[:n_owned]
        order = torch.argsort(own, stable=True)
        counts = torch.bincount(own, minlength=n_agg)
        ptr = torch.zeros(n_agg + 1, dtype=torch.int32, device=x.device)
        ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        self.n_agg, self.agg = n_agg, agg
        self.node_agg = agg.to(torch.int32).contiguous()
        self.agg_ptr, self.agg_nodes = ptr, order.to(torch.int32).contiguous()
        fx = None if fixed is None else fixed.to(torch.uint8).contiguous()
        cen = None if centers is None else centers.to(torch.float64).contiguous()
        self._setup_args = (x, y, z, self.node_agg, self.agg_ptr, self.agg_nodes, fx, cen)
        self._pattern, self._n_owned, self._allreduce = (rowptr, colidx), n_owned, allreduce
        self.active = False
        self.update(vals)

    def update(self, vals):
        """New matrix values on the same mesh, constraints and aggregates (a design iteration changes the radii, not the
        lattice): only the Galerkin product and the dense factorisation are redone."""
        self._make_resident()
        self.E = self.ctx.coarse_galerkin(self._pattern[0], self._pattern[1], vals, self.n_agg, n_rows=self._n_owned)
        if self._allreduce is not None:
            self._allreduce(self.E)
        self.Einv = invert_coarse(self.E)
        if self.active:
            self.ctx.coarse_set_inverse(self.Einv)
        return self

    def _make_resident(self):
        """A context holds ONE coarse space (node tables); building another TwoLevel on the same context replaces them, so
        the tables are re-built before this one is used again."""
        if getattr(self.ctx, "_coarse_owner", None) is not self:
            self.ctx.coarse_setup(*self._setup_args)
            self.ctx._coarse_owner = self

    def __enter__(self):
        self._make_resident()
        self.ctx.coarse_set_inverse(self.Einv)
        self.active = True
        return self

    def __exit__(self, *exc):
        self.ctx.coarse_set_inverse(None)
        self.active = False
        return False

    def apply(self, r, u):
        """u += Z Einv Z^T r (one coarse correction; for tests)."""
        was = self.active
        if not was:
            self._make_resident()
            self.ctx.coarse_set_inverse(self.Einv)
        try:
            return self.ctx.coarse_apply(r, u)
        finally:
            if not was:
                self.ctx.coarse_set_inverse(None)
