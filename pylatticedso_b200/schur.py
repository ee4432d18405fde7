"""Per-cell Schur complements on the GPU: drop-in for
``pyLatticeSim.utils_schur.get_schur_complement`` (utils_schur.py:22-53) and the
batched form that feeds ``LatticeSim.calculate_schur_complement_cells``
(lattice_sim.py:846-919) and the npz dataset schema (utils_schur.py:55-72).
"""
from __future__ import annotations

import numpy as np

from . import lib as L
from .fem import KAPPA, material_constants
from .mesh import BeamMesh, flatten_lattice, mesh_from_synthetic


def local_cell_mesh(mesh: BeamMesh, boundary_nodes):
    """Renumber a one-cell mesh for ``lat_schur_batch``: the boundary nodes first, in the given
    order (``cell.node_in_order_simulation``), then the interior nodes with the strut-interior
    (degree-2) nodes before the interior joints -- the order in which the partial Cholesky
    produces the least fill.  Returns (perm_old_to_new, xyz_local[nn,3], len0, len1)."""
    nn = mesh.n_nodes
    boundary_nodes = np.asarray(boundary_nodes, dtype=np.int64)
    is_b = np.zeros(nn, dtype=bool)
    is_b[boundary_nodes] = True
    deg = np.bincount(np.concatenate([mesh.en0, mesh.en1]), minlength=nn)
    interior = np.flatnonzero(~is_b)
    chain = interior[(interior >= mesh.n_points)]          # strut-interior nodes, already beam-major
    joints = interior[(interior < mesh.n_points)]
    joints = joints[np.argsort(deg[joints], kind="stable")]
    order = np.concatenate([boundary_nodes, chain, joints])
    perm = np.empty(nn, dtype=np.int64)
    perm[order] = np.arange(nn)
    xyz = mesh.xyz[order]
    return perm, xyz, perm[mesh.en0].astype(np.int32), perm[mesh.en1].astype(np.int32)


def strut_chains(xyz, len0, len1, n_bnd_nodes, tol=1e-9, allow_trivial=False):
    """Straight element chains of a local cell mesh (boundary nodes = the first ``n_bnd_nodes`` local nodes).

    A chain node is an interior node with exactly two incident elements that are collinear; every maximal run of
    chain nodes between two joints is one chain.  Returns host arrays for ``lat_schur_batch_chains``:
    ptr / elem / flip (elements of each chain in walking order), a / b (end joints in the REDUCED numbering: the
    boundary nodes keep their index, interior joints follow in ascending local order) and n_joints -- or ``None``
    when nothing can be condensed (every strut is a single element; ``allow_trivial=True`` then returns one
    single-element chain per element, which is what the star-cell kernel wants)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    len0, len1 = np.asarray(len0, dtype=np.int64), np.asarray(len1, dtype=np.int64)
    nn, ne = xyz.shape[0], len0.shape[0]
    deg = np.bincount(np.concatenate([len0, len1]), minlength=nn)
    d = xyz[len1] - xyz[len0]
    d /= np.linalg.norm(d, axis=1)[:, None]
    incident = [[] for _ in range(nn)]
    for e in range(ne):
        incident[len0[e]].append(e)
        incident[len1[e]].append(e)
    is_chain = np.zeros(nn, dtype=bool)
    for n in range(n_bnd_nodes, nn):
        if deg[n] == 2:
            e0, e1 = incident[n]
            is_chain[n] = np.linalg.norm(np.cross(d[e0], d[e1])) < tol
    if not is_chain.any() and not allow_trivial:
        return None
    joints = np.flatnonzero(~is_chain)
    red = np.full(nn, -1, dtype=np.int64)
    red[joints] = np.arange(joints.size)          # boundary nodes are the first local nodes -> they stay first
    used = np.zeros(ne, dtype=bool)
    ptr, elem, flip, ca, cb = [0], [], [], [], []
    for e0 in range(ne):
        if used[e0] or (is_chain[len0[e0]] and is_chain[len1[e0]]):
            continue
        start, nxt, fl = (len0[e0], len1[e0], 0) if not is_chain[len0[e0]] else (len1[e0], len0[e0], 1)
        used[e0] = True
        elem.append(e0); flip.append(fl)
        while is_chain[nxt]:
            e = [q for q in incident[nxt] if not used[q]][0]
            used[e] = True
            fl = 0 if len0[e] == nxt else 1
            elem.append(e); flip.append(fl)
            nxt = len1[e] if fl == 0 else len0[e]
        ptr.append(len(elem)); ca.append(red[start]); cb.append(red[nxt])
    if not used.all():
        return None                                # a closed ring of chain nodes: leave it to the dense path
    i32 = lambda v: np.asarray(v, dtype=np.int32)
    return dict(ptr=i32(ptr), elem=i32(elem), flip=i32(flip), a=i32(ca), b=i32(cb), n_joints=int(joints.size))


def is_star(chains, n_bnd_nodes):
    """One interior joint joined to every boundary joint by exactly one strut (a BCC cell at any subdivision)."""
    if chains is None or chains["n_joints"] != n_bnd_nodes + 1 or len(chains["a"]) != n_bnd_nodes or n_bnd_nodes > 16:
        return False
    c = n_bnd_nodes
    corners = [int(b) if int(a) == c else (int(a) if int(b) == c else -1) for a, b in zip(chains["a"], chains["b"])]
    return all(0 <= k < c for k in corners) and len(set(corners)) == c


def chain_groups(chains, elem_group):
    """Radius group of every chain, or None when the elements of some chain do not share one group."""
    eg = np.asarray(elem_group, dtype=np.int64)
    out = np.empty(len(chains["a"]), dtype=np.int32)
    for k in range(len(chains["a"])):
        g = eg[chains["elem"][chains["ptr"][k]: chains["ptr"][k + 1]]]
        if (g != g[0]).any():
            return None
        out[k] = g[0]
    return out


def _chains_to_device(chains, dev):
    import torch
    out = {k: torch.from_numpy(v).to(dev) for k, v in chains.items() if k != "n_joints"}
    out["n_joints"] = chains["n_joints"]
    return out


def get_schur_complement(lattice, cell_index=None, elements_per_strut="gmsh", ctx=None):
    """Drop-in for ``get_schur_complement(lattice, cell_index) -> ndarray[nB, nB]`` (C order;
    boundary DOF order = 6 DOFs of each node of ``cell.node_in_order_simulation``)."""
    import torch
    if cell_index is None and lattice.get_number_cells() > 1:
        raise ValueError("The lattice must contain only one cell for Schur complement calculation or specify a "
                         "cell_index.")                       # utils_schur.py:35-36
    cell = lattice.cells[0] if cell_index is None else lattice.cells[cell_index]
    cell.define_node_order_to_simulate()                      # utils_schur.py:39
    mesh = flatten_lattice(lattice, cell.index, elements_per_strut)
    loc = {int(i): k for k, i in enumerate(mesh.point_index)}
    bnd = np.array([loc[p.index] for p in cell.node_in_order_simulation], dtype=np.int64)
    E, nu = material_constants(lattice)
    ctx = ctx or L.default_context()
    perm, xyz, l0, l1 = local_cell_mesh(mesh, bnd)
    dev = ctx.device
    chains = strut_chains(xyz, l0, l1, len(bnd))
    if chains is None:
        trivial = strut_chains(xyz, l0, l1, len(bnd), allow_trivial=True)
        chains = trivial if is_star(trivial, len(bnd)) else None
    args = (torch.from_numpy(xyz[None]).to(dev), torch.from_numpy(l0).to(dev), torch.from_numpy(l1).to(dev),
            torch.from_numpy(mesh.rad[None].copy()).to(dev))
    if chains is not None:      # strut pre-pass: condense the ~18 elements of every strut first (star cells: warp kernel)
        S = ctx.schur_batch_struts(*args, _chains_to_device(chains, dev), len(bnd), E, nu, KAPPA)
    else:
        S = ctx.schur_batch(*args, len(bnd), E, nu, KAPPA)
    out = S[0].cpu().numpy()
    if not np.isfinite(out).all():
        raise RuntimeError("Schur complement: interior stiffness block is not positive definite")
    return out


def schur_gradients(lattice, cell, radii_params, elements_per_strut="gmsh", ctx=None):
    """Drop-in for ``LatticeSim._compute_schur_gradients(cell, radii_params) -> [dS/dr_j]`` (lattice_sim.py:1020-1054).
    The reference differentiates ``get_schur_complement`` by central finite differences (2 n_geom extra condensations,
    ~1e-6 relative noise); here dS_j = E^T (dK/dr_j) E is evaluated analytically by ``lat_schur_batch`` in the same
    launch that condenses the cell.  Element radius = ``radii_params[beam.type_beam]`` times the penalisation factor
    of the beam (x1.5 on ``beam_mod`` segments, beam.py:405-436), exactly what ``change_beam_radius`` would set."""
    import torch
    cell.define_node_order_to_simulate()
    mesh = flatten_lattice(lattice, cell.index, elements_per_strut)
    loc = {int(i): k for k, i in enumerate(mesh.point_index)}
    bnd = np.array([loc[p.index] for p in cell.node_in_order_simulation], dtype=np.int64)
    E, nu = material_constants(lattice)
    ctx = ctx or L.default_context()
    perm, xyz, l0, l1 = local_cell_mesh(mesh, bnd)
    radii_params = [float(r) for r in radii_params]
    types = np.asarray(mesh.type_of_elem, dtype=np.int64)
    if types.max(initial=0) >= len(radii_params):
        raise ValueError("radii_params has fewer entries than the cell has beam types")
    rad = np.asarray(radii_params)[types] * mesh.chain
    dev = ctx.device
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
    chains = strut_chains(xyz, l0, l1, len(bnd), allow_trivial=True)
    fast = chains is not None and (is_star(chains, len(bnd)) or chains["n_joints"] == len(bnd))
    cg = chain_groups(chains, types) if fast else None
    if cg is not None:          # star cell / no interior joint: sensitivities through the differentiated strut pre-pass
        S, dS = ctx.schur_batch_struts(t(xyz[None], np.float64), t(l0, np.int32), t(l1, np.int32), t(rad[None], np.float64),
                                       _chains_to_device(chains, dev), len(bnd), E, nu, KAPPA, chain_group=t(cg, np.int32),
                                       drad_chain=t(mesh.chain, np.float64), n_grad=len(radii_params))
    else:
        S, dS = ctx.schur_batch(t(xyz[None], np.float64), t(l0, np.int32), t(l1, np.int32), t(rad[None], np.float64), len(bnd),
                                E, nu, KAPPA, elem_group=t(types, np.int32), chain=t(mesh.chain, np.float64),
                                n_grad=len(radii_params))
    out = dS[0].cpu().numpy()
    if not np.isfinite(out).all():
        raise RuntimeError("Schur sensitivities: interior stiffness block is not positive definite")
    return [np.ascontiguousarray(out[j]) for j in range(len(radii_params))]


class CellBatch:
    """Topology shared by a batch of cells + per-cell coordinates/radii resident on the GPU."""

    def __init__(self, ctx, xyz, len0, len1, rad, n_bnd_nodes, young, nu, kappa=KAPPA, elem_group=None, chain=None,
                 n_grad=0):
        import torch
        self.ctx = ctx
        dev = ctx.device
        t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
        self.xyz, self.len0, self.len1, self.rad = t(xyz, np.float64), t(len0, np.int32), t(len1, np.int32), t(rad, np.float64)
        self.n_bnd_nodes = int(n_bnd_nodes)
        self.young, self.nu, self.kappa = young, nu, kappa
        self.elem_group = None if elem_group is None else t(elem_group, np.int32)
        self.chain = None if chain is None else t(chain, np.float64)
        self.n_grad = int(n_grad)
        # strut pre-pass: topology is shared, so the chains are found once on cell 0; star cells (one interior joint,
        # BCC) also take single-element "chains" and carry the sensitivities through the pre-pass
        ch = strut_chains(np.asarray(xyz)[0], len0, len1, self.n_bnd_nodes)
        if ch is None:
            # single-element struts: star cells and cells without an interior joint (Octet) still go through the
            # strut path (half-warp star kernel / direct assembly of the joint-only matrix)
            tr = strut_chains(np.asarray(xyz)[0], len0, len1, self.n_bnd_nodes, allow_trivial=True)
            ch = tr if (tr is not None and (is_star(tr, self.n_bnd_nodes) or tr["n_joints"] == self.n_bnd_nodes)) else None
        self.star = is_star(ch, self.n_bnd_nodes)
        self.direct = ch is not None and ch["n_joints"] == self.n_bnd_nodes      # no interior joint: S is assembled directly
        self.chains = None if ch is None else _chains_to_device(ch, dev)
        self.chain_group = None
        if (self.star or self.direct) and elem_group is not None:
            cg = chain_groups(ch, elem_group)
            self.chain_group = None if cg is None else t(cg, np.int32)

    def schur(self, with_gradients=False, use_chains=True):
        if use_chains and self.chains is not None and not with_gradients:
            return self.ctx.schur_batch_struts(self.xyz, self.len0, self.len1, self.rad, self.chains, self.n_bnd_nodes,
                                               self.young, self.nu, self.kappa)
        if use_chains and with_gradients and (self.star or self.direct) and self.chain_group is not None:
            return self.ctx.schur_batch_struts(self.xyz, self.len0, self.len1, self.rad, self.chains, self.n_bnd_nodes,
                                               self.young, self.nu, self.kappa, chain_group=self.chain_group,
                                               drad_chain=self.chain, n_grad=self.n_grad)
        if with_gradients:
            return self.ctx.schur_batch(self.xyz, self.len0, self.len1, self.rad, self.n_bnd_nodes, self.young, self.nu,
                                        self.kappa, self.elem_group, self.chain, self.n_grad)
        return self.ctx.schur_batch(self.xyz, self.len0, self.len1, self.rad, self.n_bnd_nodes, self.young, self.nu,
                                    self.kappa)


def bcc_cell_order_nodes(pxyz, box):
    """``Cell.define_node_order_to_simulate`` (cell.py:611-680) for arrays: nodes on the cell box, each
    assigned to the first face of (Xmin, Xmax, Ymin, Ymax, Zmin, Zmax) it lies on, sorted in-plane."""
    x0, x1, y0, y1, z0, z1 = box
    tol = 1e-9
    faces = [np.abs(pxyz[:, 0] - x0) <= tol, np.abs(pxyz[:, 0] - x1) <= tol, np.abs(pxyz[:, 1] - y0) <= tol,
             np.abs(pxyz[:, 1] - y1) <= tol, np.abs(pxyz[:, 2] - z0) <= tol, np.abs(pxyz[:, 2] - z1) <= tol]
    taken = np.zeros(pxyz.shape[0], dtype=bool)
    order = []
    keys = [(1, 2, 0), (1, 2, 0), (0, 2, 1), (0, 2, 1), (0, 1, 2), (0, 1, 2)]
    for f, key in zip(faces, keys):
        idx = np.flatnonzero(f & ~taken)
        taken[idx] = True
        srt = np.lexsort((pxyz[idx, key[2]], pxyz[idx, key[1]], pxyz[idx, key[0]]))
        order.extend(idx[srt].tolist())
    return np.array(order, dtype=np.int64)


def synthetic_cell_batch(ctx, geom, radii, elements_per_strut, young, nu, cell_size=1.0, with_gradients=False):
    """One unit cell of ``geom`` per entry of ``radii`` (shape [n_cells]) -- BASELINE config 4."""
    from .mesh import synthetic_lattice
    lat = synthetic_lattice(geom, (1, 1, 1), [1.0], cell_size=(cell_size,) * 3)
    mesh = mesh_from_synthetic(lat, elements_per_strut)
    bnd = bcc_cell_order_nodes(lat.pxyz, (0, cell_size, 0, cell_size, 0, cell_size))
    perm, xyz, l0, l1 = local_cell_mesh(mesh, bnd)
    radii = np.asarray(radii, dtype=np.float64)
    n_cells = radii.shape[0]
    xyz_b = np.broadcast_to(xyz[None], (n_cells,) + xyz.shape).copy()
    rad_b = np.repeat(radii[:, None], mesh.n_elems, axis=1)
    grp = np.zeros(mesh.n_elems, dtype=np.int32) if with_gradients else None
    ch = np.ones(mesh.n_elems) if with_gradients else None
    return CellBatch(ctx, xyz_b, l0, l1, rad_b, len(bnd), young, nu, elem_group=grp, chain=ch,
                     n_grad=1 if with_gradients else 0), bnd


def save_schur_dataset(path, radius_values, schur_matrices):
    """Write a batch of Schur complements in the reference's dataset schema (utils_schur.py:55-72:
    ``radius_values`` [n, n_geom], ``schur_matrices`` [n, nB, nB]) so that
    ``load_schur_complement_dataset`` (utils_schur.py:93-129), the preconditioner approximations
    (lattice_sim.py:1312-1329) and the greedy reduced basis can consume GPU-computed batches."""
    import torch
    S = schur_matrices.cpu().numpy() if torch.is_tensor(schur_matrices) else np.asarray(schur_matrices)
    rv = np.asarray(radius_values, dtype=np.float64)
    if rv.ndim == 1:
        rv = rv[:, None]
    if rv.shape[0] != S.shape[0]:
        raise ValueError("radius_values and schur_matrices must have the same leading dimension")
    np.savez(path, radius_values=rv, schur_matrices=S)
    return path


def load_schur_dataset(path):
    """{tuple(radii): S} exactly like ``load_schur_complement_dataset`` (utils_schur.py:108-124)."""
    data = np.load(path, allow_pickle=True)
    rv, S = data["radius_values"], data["schur_matrices"]
    if np.ndim(rv[0]) == 0 or np.ndim(rv) == 1:
        return {tuple(np.atleast_1d(rv)): S}
    return {tuple(r): m for r, m in zip(rv, S)}
