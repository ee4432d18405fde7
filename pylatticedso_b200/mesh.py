"""Host-side flattening of a lattice into the SoA arrays the CUDA path consumes.

Two producers, one layout:

* :func:`flatten_lattice` walks an (unmodified) pyLatticeDSO ``Lattice`` /
  ``LatticeSim`` / ``LatticeOpti`` object graph -- anything exposing
  ``cells[*].points_cell / beams_cell``, ``Point.x,y,z,index`` and
  ``Beam.point1,point2,radius,index,beam_mod`` -- exactly like the reference's
  gmsh front end does (``pyLatticeSim/lattice_generation.py:105-175``).
* :func:`synthetic_lattice` builds the same arrays for regular BCC / Octet /
  user-table lattices with numpy only, without creating Python ``Beam``
  objects (the reference costs ~65 us per beam), in the
  reference's own numbering: ``node.index`` = rank in the (x, y, z) sort,
  ``beam.index`` = rank in the (min end, max end, radius) sort
  (``pyLatticeDesign/lattice.py:665-698``), cells i-major
  (``lattice.py:448-453``), shared beams keep the radius of the first cell that
  created them (``cell.py:366-378``).

Canonical FE numbering (the only observable numbering):
lattice points first, in ``node.index`` order; strut-interior nodes appended
beam-major (``beam.index`` order), from ``point1`` to ``point2``; per-node DOF
order [ux, uy, uz, rx, ry, rz] (``point.py:68``); global DOF = 6*node + d.
Elements are beam-major, ``point1`` -> ``point2``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

NDOF = 6
MESH_FRACTION = 0.05  # lattice_generation.py:50-64: h = 0.05 * cell_size_x

# Unit-cell strut tables (fractional coordinates), pyLatticeDesign/geometries/{BCC,Octet}.json
_BCC = np.array([
    [0.0, 0.0, 0.0, 0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 1.0, 1.0, 1.0],
    [0.5, 0.5, 0.5, 1.0, 1.0, 0.0], [0.5, 0.5, 0.5, 0.0, 0.0, 1.0],
    [0.5, 0.5, 0.5, 0.0, 1.0, 0.0], [0.5, 0.5, 0.5, 0.0, 1.0, 1.0],
    [1.0, 0.0, 1.0, 0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 1.0, 0.0, 0.0]])


# Octet-truss unit cell: 24 corner -> face-centre struts + 12 octahedron struts,
# row order and orientation as in the reference's geometry table (data, not code).
_OCTET = np.array([
    [0.0, 0.0, 0.0, 0.5, 0.0, 0.5],
    [1.0, 0.0, 1.0, 0.5, 0.0, 0.5],
    [0.0, 0.0, 1.0, 0.5, 0.0, 0.5],
    [1.0, 0.0, 0.0, 0.5, 0.0, 0.5],
    [0.0, 0.0, 0.0, 0.0, 0.5, 0.5],
    [0.0, 1.0, 1.0, 0.0, 0.5, 0.5],
    [0.0, 0.0, 1.0, 0.0, 0.5, 0.5],
    [0.0, 1.0, 0.0, 0.0, 0.5, 0.5],
    [0.0, 0.0, 0.0, 0.5, 0.5, 0.0],
    [1.0, 1.0, 0.0, 0.5, 0.5, 0.0],
    [1.0, 0.0, 0.0, 0.5, 0.5, 0.0],
    [0.0, 1.0, 0.0, 0.5, 0.5, 0.0],
    [0.0, 0.0, 1.0, 0.5, 0.5, 1.0],
    [1.0, 1.0, 1.0, 0.5, 0.5, 1.0],
    [1.0, 0.0, 1.0, 0.5, 0.5, 1.0],
    [0.0, 1.0, 1.0, 0.5, 0.5, 1.0],
    [1.0, 0.5, 0.5, 1.0, 1.0, 1.0],
    [1.0, 0.0, 0.0, 1.0, 0.5, 0.5],
    [1.0, 0.5, 0.5, 1.0, 1.0, 0.0],
    [1.0, 0.0, 1.0, 1.0, 0.5, 0.5],
    [0.5, 1.0, 0.5, 1.0, 1.0, 1.0],
    [0.0, 1.0, 0.0, 0.5, 1.0, 0.5],
    [0.5, 1.0, 0.5, 1.0, 1.0, 0.0],
    [0.0, 1.0, 1.0, 0.5, 1.0, 0.5],
    [0.5, 0.0, 0.5, 0.5, 0.5, 0.0],
    [0.5, 0.0, 0.5, 0.0, 0.5, 0.5],
    [0.5, 0.0, 0.5, 1.0, 0.5, 0.5],
    [0.5, 0.0, 0.5, 0.5, 0.5, 1.0],
    [0.5, 1.0, 0.5, 0.5, 0.5, 0.0],
    [0.5, 1.0, 0.5, 0.0, 0.5, 0.5],
    [0.5, 1.0, 0.5, 1.0, 0.5, 0.5],
    [0.5, 1.0, 0.5, 0.5, 0.5, 1.0],
    [1.0, 0.5, 0.5, 0.5, 0.5, 0.0],
    [0.5, 0.5, 0.0, 0.0, 0.5, 0.5],
    [0.5, 0.5, 1.0, 0.0, 0.5, 0.5],
    [0.5, 0.5, 1.0, 1.0, 0.5, 0.5]])


GEOMETRY_TABLES = {"BCC": _BCC, "Octet": _OCTET}


@dataclass
class BeamMesh:
    """SoA finite-element mesh of 2-node beam elements (host copy)."""
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    en0: np.ndarray            # int32 [E] first node of each element
    en1: np.ndarray            # int32 [E]
    rad: np.ndarray            # float64 [E] element radius (x1.5 already applied on penalised beams)
    beam_of_elem: np.ndarray   # int64 [E] beam.index of the owning beam
    chain: np.ndarray          # float64 [E] d(rad_e)/d(base radius): 1.5 on beam_mod beams else 1
    n_points: int              # the first n_points nodes are lattice Points
    point_index: np.ndarray    # int64 [n_points] node.index of those points
    cell_of_elem: np.ndarray | None = None   # int64 [E] owning cell (first owner) or None
    type_of_elem: np.ndarray | None = None   # int64 [E] beam.type_beam (geometry slot) or None
    meta: dict = field(default_factory=dict)

    @property
    def n_nodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def n_elems(self) -> int:
        return int(self.en0.shape[0])

    @property
    def n_dof(self) -> int:
        return NDOF * self.n_nodes

    @property
    def xyz(self) -> np.ndarray:
        return np.stack([self.x, self.y, self.z], axis=1)


def gmsh_segments(length, h):
    """Elements per strut produced by the reference mesh: max(1, int(L/h + 0.99))
    (``lattice_generation.py:50-64,119,161``; pinned by the 30 Schur goldens)."""
    length = np.asarray(length, dtype=np.float64)
    return np.maximum(1, (length / h + 0.99).astype(np.int64))


def _subdivide(pxyz, b_p1, b_p2, b_rad, b_idx, b_chain, nseg, n_points, point_index,
               b_cell=None, b_type=None, meta=None) -> BeamMesh:
    """Split every beam into ``nseg[b]`` uniform elements; interior nodes are
    appended beam-major after the ``n_points`` lattice points."""
    nb = b_p1.shape[0]
    nseg = np.asarray(nseg, dtype=np.int64)
    if nb and int(nseg.max()) == 1:                      # one element per strut: the beams ARE the elements
        cp = lambda v, d: np.array(v, dtype=d, copy=True)
        return BeamMesh(
            x=np.ascontiguousarray(pxyz[:, 0]), y=np.ascontiguousarray(pxyz[:, 1]), z=np.ascontiguousarray(pxyz[:, 2]),
            en0=b_p1.astype(np.int32), en1=b_p2.astype(np.int32), rad=cp(b_rad, np.float64), beam_of_elem=cp(b_idx, np.int64),
            chain=cp(b_chain, np.float64), n_points=int(n_points), point_index=np.asarray(point_index, dtype=np.int64),
            cell_of_elem=None if b_cell is None else cp(b_cell, np.int64),
            type_of_elem=None if b_type is None else cp(b_type, np.int64), meta=meta or {})
    a = pxyz[b_p1]
    c = pxyz[b_p2]
    n_int = nseg - 1
    int_off = np.concatenate([[0], np.cumsum(n_int)])  # per-beam offset into interior nodes
    tot_int = int(int_off[-1])
    el_off = np.concatenate([[0], np.cumsum(nseg)])
    tot_el = int(el_off[-1])
    # interior node coordinates: a + (c - a) * (k / nseg), k = 1..nseg-1
    bi = np.repeat(np.arange(nb), n_int)
    k = np.arange(tot_int) - int_off[bi] + 1
    frac = k / nseg[bi]
    ixyz = a[bi] + (c[bi] - a[bi]) * frac[:, None]
    xyz = np.concatenate([pxyz, ixyz], axis=0) if tot_int else pxyz.copy()
    # elements
    be = np.repeat(np.arange(nb), nseg)
    j = np.arange(tot_el) - el_off[be]           # segment number within the beam
    first = j == 0
    last = j == nseg[be] - 1
    int_base = n_points + int_off[be]
    en0 = np.where(first, b_p1[be], int_base + j - 1)
    en1 = np.where(last, b_p2[be], int_base + j)
    return BeamMesh(
        x=np.ascontiguousarray(xyz[:, 0]), y=np.ascontiguousarray(xyz[:, 1]),
        z=np.ascontiguousarray(xyz[:, 2]),
        en0=en0.astype(np.int32), en1=en1.astype(np.int32),
        rad=np.asarray(b_rad, dtype=np.float64)[be].copy(),
        beam_of_elem=np.asarray(b_idx, dtype=np.int64)[be].copy(),
        chain=np.asarray(b_chain, dtype=np.float64)[be].copy(),
        n_points=int(n_points), point_index=np.asarray(point_index, dtype=np.int64),
        cell_of_elem=None if b_cell is None else np.asarray(b_cell, dtype=np.int64)[be].copy(),
        type_of_elem=None if b_type is None else np.asarray(b_type, dtype=np.int64)[be].copy(),
        meta=meta or {})


def flatten_lattice(lattice, cell_index=None, elements_per_strut="gmsh") -> BeamMesh:
    """Flatten a pyLatticeDSO lattice object graph (or one of its cells).

    Mirrors what ``latticeGeneration.generate_nodes/generate_beams`` feed to
    gmsh (``lattice_generation.py:105-175``): every node of the selected cells,
    every beam with ``radius > 0`` (:158) once, uniform subdivision.  With
    ``cell_index`` only that cell's nodes/beams are taken (``utils_schur.py:46``).
    ``elements_per_strut``: ``"gmsh"`` (reference mesh rule) or an int.
    """
    cells = list(lattice.cells) if cell_index is None else [c for c in lattice.cells if c.index == cell_index]
    if not cells:
        raise ValueError(f"no cell with index {cell_index}")
    pts, beams, first_cell = {}, {}, {}
    for c in cells:
        for p in c.points_cell:
            pts[p.index] = p
        for b in c.beams_cell:
            if b.radius > 0 and b.index not in beams:
                beams[b.index] = b
                first_cell[b.index] = c.index
    order = sorted(pts)
    loc = {idx: k for k, idx in enumerate(order)}
    pxyz = np.array([[pts[i].x, pts[i].y, pts[i].z] for i in order], dtype=np.float64).reshape(-1, 3)
    bord = sorted(beams)
    bl = [beams[i] for i in bord]
    b_p1 = np.array([loc[b.point1.index] for b in bl], dtype=np.int64)
    b_p2 = np.array([loc[b.point2.index] for b in bl], dtype=np.int64)
    b_rad = np.array([b.radius for b in bl], dtype=np.float64)
    b_mod = np.array([bool(getattr(b, "beam_mod", False)) for b in bl], dtype=bool)
    coef = np.array([getattr(b, "penalization_coefficient", 1.5) for b in bl], dtype=np.float64)
    b_chain = np.where(b_mod, coef, 1.0)
    b_type = np.array([getattr(b, "type_beam", 0) for b in bl], dtype=np.int64)
    b_cell = np.array([first_cell[i] for i in bord], dtype=np.int64)
    if elements_per_strut == "gmsh":
        h = MESH_FRACTION * lattice.cell_size_x
        L = np.linalg.norm(pxyz[b_p2] - pxyz[b_p1], axis=1)
        nseg = gmsh_segments(L, h)
    else:
        nseg = np.full(len(bl), int(elements_per_strut), dtype=np.int64)
    mesh = _subdivide(pxyz, b_p1, b_p2, b_rad, np.array(bord, dtype=np.int64), b_chain, nseg,
                      len(order), np.array(order, dtype=np.int64), b_cell=b_cell, b_type=b_type,
                      meta={"source": "object_graph", "cell_index": cell_index})
    mesh.meta["points"] = [pts[i] for i in order]
    return mesh


def cell_boundary_dofs(cell, mesh: BeamMesh) -> np.ndarray:
    """Boundary DOF list of a cell in the reference's Schur ordering: 6 DOFs of
    each node of ``cell.node_in_order_simulation`` (``utils_schur.py:39-44``,
    ``cell.py:611-680``)."""
    if getattr(cell, "node_in_order_simulation", None) is None:
        cell.define_node_order_to_simulate()
    loc = {int(i): k for k, i in enumerate(mesh.point_index)}
    nodes = np.array([loc[p.index] for p in cell.node_in_order_simulation], dtype=np.int64)
    return (nodes[:, None] * NDOF + np.arange(NDOF)[None, :]).ravel()


# --------------------------------------------------------------------------
# vectorised regular-lattice generator
# --------------------------------------------------------------------------
@dataclass
class SyntheticLattice:
    """Array form of a regular lattice in the reference's numbering."""
    pxyz: np.ndarray        # [Np,3] node coordinates, row = node.index
    b_p1: np.ndarray        # [Nb] node.index of Beam.point1, row = beam.index
    b_p2: np.ndarray
    b_rad: np.ndarray       # [Nb]
    b_cell: np.ndarray      # [Nb] index of the first cell that created the beam
    b_type: np.ndarray      # [Nb] geometry slot
    n_cells: tuple
    cell_size: tuple
    cell_radii: np.ndarray  # [Nc, n_geom] radius actually used by each cell (after grading)
    geom_types: tuple

    @property
    def n_cells_total(self):
        return int(np.prod(self.n_cells))

    def cell_index(self, i, j, k):
        return (i * self.n_cells[1] + j) * self.n_cells[2] + k


def _grid_lattice(tables, ci, cj, ck, gcell, org, size, cr, dims, i_lo, i_hi):
    """Sort-free numbering for tables whose fractional coordinates are multiples of 1/2 (BCC, Octet, ...).

    On the half-cell grid every strut end has integer coordinates, so the reference's numbering rules become dense
    scatters + prefix sums: node.index = rank of the grid key (= rank in the (x,y,z) sort, cell.py:317-321); a strut is
    (lower end, offset to the upper end) and the lexicographic order of the offsets IS the order of the upper ends, so
    the occupied slots of a dense [node][offset] array are the struts in beam.index order (lattice.py:675-683); "first
    creation wins" (cell.py:366-378) = the lowest creating instance, found with a reversed scatter.  Bit-identical to
    the generic sort/unique path (tests/test_mesh.py) at a fraction of its time: 72 M strut ends for Octet 100^3."""
    nx, ny, nz = dims
    T = np.concatenate(tables, axis=0)
    geom_of_row = np.concatenate([np.full(t.shape[0], g, dtype=np.int64) for g, t in enumerate(tables)])
    T2 = np.rint(T * 2.0).astype(np.int64)                       # (nbt, 6) half-cell units
    nbt = T2.shape[0]
    ncl = ci.shape[0]
    Gy, Gz = 2 * ny + 1, 2 * nz + 1
    n_slots = (2 * (i_hi - i_lo) + 1) * Gy * Gz
    cbase = ((2 * (ci - i_lo)) * Gy + 2 * cj) * Gz + 2 * ck       # key of the cell origin
    off1 = (T2[:, 0] * Gy + T2[:, 1]) * Gz + T2[:, 2]
    off2 = (T2[:, 3] * Gy + T2[:, 4]) * Gz + T2[:, 5]
    key1 = (cbase[:, None] + off1[None, :]).ravel()
    key2 = (cbase[:, None] + off2[None, :]).ravel()
    n_inst = key1.shape[0]
    # first creation of every grid point: ends 1 of all instances are created before ends 2 (allp = [e1; e2])
    first = np.full(n_slots, -1, dtype=np.int64)
    first[key2[::-1]] = np.arange(2 * n_inst - 1, n_inst - 1, -1)
    first[key1[::-1]] = np.arange(n_inst - 1, -1, -1)           # repeated index: the last assignment wins = lowest instance
    occ = first >= 0
    node_of_slot = np.cumsum(occ, dtype=np.int64) - 1
    slots = np.flatnonzero(occ)
    npnt = slots.shape[0]
    fidx = first[slots]
    del first, occ
    end2 = fidx >= n_inst
    inst = np.where(end2, fidx - n_inst, fidx)
    cell_l, row = inst // nbt, inst % nbt
    frac = np.where(end2[:, None], T[row, 3:6], T[row, 0:3])
    pxyz = frac * size[None, :] + org[cell_l]                   # the same expression as the generic path
    n1 = node_of_slot[key1]
    n2 = node_of_slot[key2]
    del node_of_slot
    # struts: canonical offset (upper end - lower end) per table row; node order == key order
    swap = off2 < off1
    d = np.where(swap[:, None], T2[:, 0:3] - T2[:, 3:6], T2[:, 3:6] - T2[:, 0:3])
    dkey = (d[:, 0] * (4 * Gy) + d[:, 1]) * (4 * Gz) + d[:, 2]  # lexicographic in (dx, dy, dz); |d| <= 2
    codes_u, code = np.unique(dkey, return_inverse=True)
    nd = codes_u.shape[0]
    code = code.ravel().astype(np.int64)
    lo = np.where(np.tile(swap, ncl), n2, n1)
    bkey = lo * nd + np.tile(code, ncl)
    del lo
    bfirst = np.full(npnt * nd, -1, dtype=np.int64)
    bfirst[bkey[::-1]] = np.arange(n_inst - 1, -1, -1)
    del bkey
    binst = bfirst[np.flatnonzero(bfirst >= 0)]                 # already in (lower end, upper end) order
    del bfirst
    bcell_l, brow = binst // nbt, binst % nbt
    return (pxyz, n1[binst], n2[binst], cr[bcell_l, geom_of_row[brow]], gcell[bcell_l], geom_of_row[brow])


def _grid_lattice_torch(tables, ci, cj, ck, gcell, org, size, cr, dims, i_lo, i_hi, device):
    """:func:`_grid_lattice` with the large scatters, prefix sums and gathers as torch ops on ``device`` (the GPU of the
    rank that generates its slab): "first creation wins" = the LOWEST creating instance = ``scatter_reduce(amin)``, which
    is deterministic; every floating-point expression is the same elementwise one.  Bit-identical outputs
    (tests/test_mesh.py compares the two on CPU tensors); returns numpy arrays like the numpy path."""
    import torch
    nx, ny, nz = dims
    T = np.concatenate(tables, axis=0)
    geom_of_row = np.concatenate([np.full(t.shape[0], g, dtype=np.int64) for g, t in enumerate(tables)])
    T2 = np.rint(T * 2.0).astype(np.int64)
    nbt, ncl = T2.shape[0], ci.shape[0]
    Gy, Gz = 2 * ny + 1, 2 * nz + 1
    n_slots = (2 * (i_hi - i_lo) + 1) * Gy * Gz
    dev = torch.device(device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cbase = t(((2 * (ci - i_lo)) * Gy + 2 * cj) * Gz + 2 * ck)
    off1 = (T2[:, 0] * Gy + T2[:, 1]) * Gz + T2[:, 2]
    off2 = (T2[:, 3] * Gy + T2[:, 4]) * Gz + T2[:, 5]
    key1 = (cbase[:, None] + t(off1)[None, :]).reshape(-1)
    key2 = (cbase[:, None] + t(off2)[None, :]).reshape(-1)
    n_inst = int(key1.shape[0])
    BIG = 2 ** 62
    # creation index: ends 1 of all instances (0 .. n_inst-1) come before ends 2 (n_inst ..): the first creation of a
    # grid point is the minimum over everything that lands on it
    first = torch.full((n_slots,), BIG, dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, torch.cat([key1, key2]), torch.arange(2 * n_inst, dtype=torch.int64, device=dev), "amin")
    occ = first < BIG
    node_of_slot = torch.cumsum(occ, 0) - 1
    fidx = first[occ]
    npnt = int(fidx.shape[0])
    del first, occ
    end2 = fidx >= n_inst
    inst = torch.where(end2, fidx - n_inst, fidx)
    cell_l, row = inst // nbt, inst % nbt
    Tt = t(T)
    frac = torch.where(end2[:, None], Tt[row, 3:6], Tt[row, 0:3])
    pxyz = frac * t(size)[None, :] + t(org)[cell_l]
    n1 = node_of_slot[key1]
    n2 = node_of_slot[key2]
    del node_of_slot, key1, key2
    swap = off2 < off1
    d = np.where(swap[:, None], T2[:, 0:3] - T2[:, 3:6], T2[:, 3:6] - T2[:, 0:3])
    dkey = (d[:, 0] * (4 * Gy) + d[:, 1]) * (4 * Gz) + d[:, 2]
    codes_u, code = np.unique(dkey, return_inverse=True)
    nd = int(codes_u.shape[0])
    code = code.ravel().astype(np.int64)
    lo = torch.where(t(swap).repeat(ncl), n2, n1)
    bkey = lo * nd + t(code).repeat(ncl)
    del lo
    bfirst = torch.full((npnt * nd,), BIG, dtype=torch.int64, device=dev)
    bfirst.scatter_reduce_(0, bkey, torch.arange(n_inst, dtype=torch.int64, device=dev), "amin")
    del bkey
    binst = bfirst[bfirst < BIG]                                   # already in (lower end, upper end) order
    del bfirst
    bcell_l, brow = binst // nbt, binst % nbt
    g_row = t(geom_of_row)[brow]
    out = (pxyz, n1[binst], n2[binst], t(cr)[bcell_l, g_row], t(gcell)[bcell_l], g_row)
    return tuple(o.cpu().numpy() for o in out)


def synthetic_lattice(geom_types, n_cells, radii, cell_size=(1.0, 1.0, 1.0),
                      grad_radius=None, cell_radii=None, i_range=None, _force_generic=False, device=None) -> SyntheticLattice:
    """Regular lattice arrays in the reference numbering.

    ``i_range = (i_lo, i_hi)`` generates only the cell layers ``i_lo <= i < i_hi`` of the
    ``n_cells`` lattice (coordinates, radii and ``b_cell`` are those of the full lattice; node and
    beam indices are ranks WITHIN the generated part, i.e. monotone in the global ones) -- the
    per-slab generator of the sharded solver (:func:`pylatticedso_b200.distributed.SlabFEM`).

    ``geom_types``: name or list of names from :data:`GEOMETRY_TABLES` (or
    ``[E,6]`` arrays of fractional strut coordinates); ``radii``: one base radius
    per geometry; ``grad_radius``: optional ``(rule, directions, parameters)`` as
    in the JSON ``gradient.radii`` block -- only ``"linear"``/``"constant"`` are
    evaluated here (``gradient_properties.py:104``); ``cell_radii``: optional
    ``[Nc, n_geom]`` explicit per-cell radii (cell-index order), overriding both.
    ``device``: a torch device (e.g. the rank's GPU) on which the half-cell-grid numbering runs
    (:func:`_grid_lattice_torch`, same result bit for bit); None = numpy on the host.
    """
    if isinstance(geom_types, str) or isinstance(geom_types, np.ndarray):
        geom_types = [geom_types]
    radii = list(np.atleast_1d(radii).astype(float))
    nx, ny, nz = (int(v) for v in n_cells)
    cs = tuple(float(v) for v in cell_size)
    nc = nx * ny * nz
    tables = [GEOMETRY_TABLES[g] if isinstance(g, str) else np.asarray(g, dtype=float) for g in geom_types]
    # cell origins: cumulative sums exactly like lattice.py:433-442 (grad_dim == 1)
    xs = np.concatenate([[0.0], np.cumsum(np.full(max(nx - 1, 0), cs[0]))])[:nx]
    ys = np.concatenate([[0.0], np.cumsum(np.full(max(ny - 1, 0), cs[1]))])[:ny]
    zs = np.concatenate([[0.0], np.cumsum(np.full(max(nz - 1, 0), cs[2]))])[:nz]
    i_lo, i_hi = (0, nx) if i_range is None else (max(0, int(i_range[0])), min(nx, int(i_range[1])))
    if i_hi <= i_lo:
        raise ValueError(f"empty cell-layer range {i_range}")
    ci, cj, ck = np.meshgrid(np.arange(i_lo, i_hi), np.arange(ny), np.arange(nz), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()            # i-major == cell.index order
    gcell = (ci * ny + cj) * nz + ck                           # cell.index in the full lattice
    org = np.stack([xs[ci], ys[cj], zs[ck]], axis=1)
    # per-cell radii
    ng = len(tables)
    if cell_radii is not None:
        cr = np.asarray(cell_radii, dtype=np.float64).reshape(nc, ng)[gcell]
    else:
        fac = np.ones(gcell.shape[0])
        if grad_radius is not None:
            rule, direction, params = grad_radius
            if rule == "linear":
                # get_grad_settings: factor table indexed by pos[d] for each direction d
                for d, idx in enumerate((ci, cj, ck)):
                    if direction[d]:
                        fac = fac * (1.0 + idx * params[d])
            elif rule != "constant":
                raise NotImplementedError(f"gradient rule {rule!r}")
        cr = np.stack([np.float64(r) * fac for r in radii], axis=1)
    size = np.array(cs)
    if not _force_generic and all(np.abs(t * 2.0 - np.rint(t * 2.0)).max() < 1e-12 and t.min() >= 0.0 and t.max() <= 1.0
                                  for t in tables):
        if device is None:
            pxyz, p1, p2, brad, bcell, btype = _grid_lattice(tables, ci, cj, ck, gcell, org, size, cr, (nx, ny, nz), i_lo, i_hi)
        else:
            pxyz, p1, p2, brad, bcell, btype = _grid_lattice_torch(tables, ci, cj, ck, gcell, org, size, cr, (nx, ny, nz),
                                                                   i_lo, i_hi, device)
        return SyntheticLattice(pxyz=pxyz, b_p1=p1.astype(np.int64), b_p2=p2.astype(np.int64), b_rad=brad,
                                b_cell=bcell.astype(np.int64), b_type=btype.astype(np.int64), n_cells=(nx, ny, nz),
                                cell_size=cs, cell_radii=cr,
                                geom_types=tuple(g if isinstance(g, str) else "custom" for g in geom_types))
    ends1, ends2, brad, bcell, btype = [], [], [], [], []
    for g, tab in enumerate(tables):
        nbt = tab.shape[0]
        # x = frac * size + origin (cell.py:340-346)
        e1 = tab[None, :, 0:3] * size[None, None, :] + org[:, None, :]
        e2 = tab[None, :, 3:6] * size[None, None, :] + org[:, None, :]
        ends1.append(e1)
        ends2.append(e2)
        ncl = gcell.shape[0]
        brad.append(np.repeat(cr[:, g], nbt).reshape(ncl, nbt))
        bcell.append(np.repeat(gcell, nbt).reshape(ncl, nbt))
        btype.append(np.full((ncl, nbt), g))
    # creation order: cell-major, then geometry slot, then table row (cell.py:293-382)
    e1 = np.concatenate(ends1, axis=1).reshape(-1, 3)
    e2 = np.concatenate(ends2, axis=1).reshape(-1, 3)
    brad = np.concatenate(brad, axis=1).ravel()
    bcell = np.concatenate(bcell, axis=1).ravel()
    btype = np.concatenate(btype, axis=1).ravel()
    allp = np.concatenate([e1, e2], axis=0)
    # node.index = rank in the (x,y,z) sort of the coordinates rounded to 9 digits (cell.py:317-321);
    # coordinates = first creation.  Lexicographic rank via per-axis ranks folded into one int64 key
    # (a 1-D unique is an order of magnitude faster than np.unique(axis=0) on 5e7 points).
    ranks, sizes = [], []
    for d in range(3):
        kd = np.round(allp[:, d], 9)
        ud, rd = np.unique(kd, return_inverse=True)
        ranks.append(rd.ravel().astype(np.int64))
        sizes.append(int(ud.shape[0]))
    flat = (ranks[0] * sizes[1] + ranks[1]) * sizes[2] + ranks[2]
    del ranks
    _, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    del flat
    inv = inv.ravel()
    pxyz = allp[first]
    n1 = inv[: e1.shape[0]]
    n2 = inv[e1.shape[0]:]
    lo = np.minimum(n1, n2)
    hi = np.maximum(n1, n2)
    npnt = pxyz.shape[0]
    pair = lo.astype(np.int64) * npnt + hi
    # dedup beams: first creation wins (cell.py:366-378)
    _, bfirst = np.unique(pair, return_index=True)
    bfirst.sort()
    n1, n2, brad, bcell, btype = n1[bfirst], n2[bfirst], brad[bfirst], bcell[bfirst], btype[bfirst]
    lo, hi = lo[bfirst], hi[bfirst]
    # beam.index = rank in (min end, max end, radius) sort (lattice.py:675-683)
    order = np.lexsort((brad, hi, lo))
    return SyntheticLattice(pxyz=pxyz, b_p1=n1[order].astype(np.int64), b_p2=n2[order].astype(np.int64),
                            b_rad=brad[order], b_cell=bcell[order].astype(np.int64),
                            b_type=btype[order].astype(np.int64), n_cells=(nx, ny, nz),
                            cell_size=cs, cell_radii=cr,
                            geom_types=tuple(g if isinstance(g, str) else "custom" for g in geom_types))


def mesh_from_synthetic(lat: SyntheticLattice, elements_per_strut=1) -> BeamMesh:
    """Subdivide a :class:`SyntheticLattice` (no joint penalisation, i.e. the
    reference with ``simulation_parameters.enable = false``)."""
    nb = lat.b_p1.shape[0]
    if elements_per_strut == "gmsh":
        h = MESH_FRACTION * lat.cell_size[0]
        L = np.linalg.norm(lat.pxyz[lat.b_p2] - lat.pxyz[lat.b_p1], axis=1)
        nseg = gmsh_segments(L, h)
    else:
        nseg = np.full(nb, int(elements_per_strut), dtype=np.int64)
    npnt = lat.pxyz.shape[0]
    return _subdivide(lat.pxyz, lat.b_p1, lat.b_p2, lat.b_rad, np.arange(nb), np.ones(nb), nseg,
                      npnt, np.arange(npnt), b_cell=lat.b_cell, b_type=lat.b_type,
                      meta={"source": "synthetic", "geom": lat.geom_types, "n_cells": lat.n_cells})


def surface_nodes(pxyz, surface: str):
    """Lattice points on a face of the lattice bounding box
    (``lattice.py:1320-1361`` for a full box: exact float ``==`` on the extremum)."""
    ax = "XYZ".index(surface[0].upper())
    v = pxyz[:, ax]
    ext = v.min() if surface.lower().endswith("min") else v.max()
    return np.flatnonzero(v == ext)


def compression_bc(mesh: BeamMesh, pxyz=None, value=-0.01):
    """BASELINE config "uniaxial compression": clamp Zmin (all 6 DOF, 0), impose
    u_z = value on Zmax.  Returns (fixed uint8[6N],
    g[6N], f[6N]) in FE numbering; only lattice points carry BCs."""
    pxyz = mesh.xyz[: mesh.n_points] if pxyz is None else pxyz
    n = mesh.n_dof
    fixed = np.zeros(n, dtype=np.uint8)
    g = np.zeros(n)
    f = np.zeros(n)
    bot = surface_nodes(pxyz, "Zmin")
    top = surface_nodes(pxyz, "Zmax")
    fixed[(bot[:, None] * NDOF + np.arange(NDOF)[None, :]).ravel()] = 1
    fixed[top * NDOF + 2] = 1
    g[top * NDOF + 2] = value
    return fixed, g, f


def bc_arrays_from_lattice(lattice, mesh: BeamMesh, dedup_point_loads=False):
    """(fixed, g, f) from ``Point.fixed_DOF / displacement_vector / applied_force``.

    Follows ``full_scale_lattice_simulation.py:39-73`` (Dirichlet) and
    ``:124-153`` (point loads): only the translational components of
    ``applied_force`` are used (:144), and -- as in the reference -- a load is
    appended once per *cell* that contains the node (:134-153 loops over
    ``cell.points_cell`` without de-duplication and the vector is assembled with
    ``ADD_VALUES`` at ``simulation_base.py:495-499``), so a node shared by k cells
    receives k times ``applied_force``.  ``dedup_point_loads=True`` applies each
    load once instead.
    """
    n = mesh.n_dof
    fixed = np.zeros(n, dtype=np.uint8)
    g = np.zeros(n)
    f = np.zeros(n)
    loc = {int(i): k for k, i in enumerate(mesh.point_index)}
    seen = set()
    for cell in lattice.cells:
        for node in cell.points_cell:
            k = loc.get(node.index)
            if k is None:
                continue
            if any(node.fixed_DOF):
                for d in range(NDOF):
                    if node.fixed_DOF[d]:
                        fixed[NDOF * k + d] = 1
                        g[NDOF * k + d] = node.displacement_vector[d]
            if any(v != 0 for v in node.applied_force):
                if dedup_point_loads and node.index in seen:
                    continue
                seen.add(node.index)
                for d in range(3):
                    f[NDOF * k + d] += float(node.applied_force[d])
    return fixed, g, f
