"""Duck-typed stand-ins for the reference's Lattice / Cell / Beam / Point objects, built from the
vectorised generator.  They expose exactly the attributes the drop-in layer reads
(pyLatticeDesign/point.py:55-72, beam.py:55-80, cell.py:86-102, lattice_sim.py:502-563), so the host
glue (flatten -> solve -> write-back) can be exercised on the GPU box, where no pyLatticeDSO checkout
exists."""
import numpy as np

from pylatticedso_b200 import mesh as M
from pylatticedso_b200.schur import bcc_cell_order_nodes


class FakePoint:
    def __init__(self, x, y, z, index):
        self.x, self.y, self.z, self.index = float(x), float(y), float(z), int(index)
        self.index_boundary = None
        self.displacement_vector = [0.0] * 6
        self.reaction_force_vector = [0.0] * 6
        self.applied_force = [0.0] * 6
        self.fixed_DOF = [False] * 6
        self.global_free_DOF_index = [None] * 6

    def __hash__(self):
        return hash((self.x, self.y, self.z))

    def __eq__(self, o):
        return isinstance(o, FakePoint) and (self.x, self.y, self.z) == (o.x, o.y, o.z)

    def set_reaction_force(self, rf):          # accumulates, point.py:372-385
        for i in range(6):
            self.reaction_force_vector[i] += rf[i]


class FakeBeam:
    def __init__(self, p1, p2, radius, index, type_beam=0):
        self.point1, self.point2, self.radius, self.index = p1, p2, float(radius), int(index)
        self.beam_mod = False
        self.penalization_coefficient = 1.5
        self.type_beam = type_beam


class FakeCell:
    def __init__(self, index, box):
        self.index, self.box = index, box
        self.points_cell, self.beams_cell = set(), set()
        self.node_in_order_simulation = None
        self.schur_complement = None

    def define_node_order_to_simulate(self):   # cell.py:611-680
        pts = [p for p in self.points_cell if p.index_boundary is not None]
        xyz = np.array([[p.x, p.y, p.z] for p in pts])
        order = bcc_cell_order_nodes(xyz, self.box)
        self.node_in_order_simulation = [pts[i] for i in order]


class FakeLattice:
    material_name = "VeroClear"

    def __init__(self, geom, n_cells, radius):
        syn = M.synthetic_lattice(geom, n_cells, [radius])
        self.syn = syn
        self.cell_size_x = 1.0
        self.points = [FakePoint(*p, k) for k, p in enumerate(syn.pxyz)]
        self.beams = [FakeBeam(self.points[a], self.points[b], r, k) for k, (a, b, r) in
                      enumerate(zip(syn.b_p1, syn.b_p2, syn.b_rad))]
        nx, ny, nz = n_cells
        self.cells = []
        for i in range(nx):
            for j in range(ny):
                for k in range(nz):
                    self.cells.append(FakeCell(len(self.cells), (i, i + 1, j, j + 1, k, k + 1)))
        for b, c in zip(self.beams, syn.b_cell):       # BCC struts belong to exactly one cell
            cell = self.cells[int(c)]
            cell.beams_cell.add(b)
            cell.points_cell.add(b.point1)
            cell.points_cell.add(b.point2)
        # index_boundary: nodes on a cell box, numbered in first-visit order (lattice_sim.py:546-563)
        counter = 0
        for cell in self.cells:
            x0, x1, y0, y1, z0, z1 = cell.box
            for p in sorted(cell.points_cell, key=lambda q: q.index):
                on = p.x in (x0, x1) or p.y in (y0, y1) or p.z in (z0, z1)
                if on and p.index_boundary is None:
                    p.index_boundary = counter
                    counter += 1
        self.max_index_boundary = counter - 1
        self.global_displacement_index = None

    def get_number_cells(self):
        return len(self.cells)

    def compression(self, value=-0.01):
        zmax = max(p.z for p in self.points)
        for p in self.points:
            if p.z == 0.0:
                p.fixed_DOF = [True] * 6
            elif p.z == zmax:
                p.fixed_DOF[2] = True
                p.displacement_vector[2] = value

    def get_global_displacement(self):          # lattice_sim.py:502-542 (withFixed=False, OnlyImposed=False)
        out, idx, seen = [], [], set()
        for cell in self.cells:
            for node in sorted(cell.points_cell, key=lambda n: (round(n.x, 9), round(n.y, 9), round(n.z, 9), n.index)):
                if node.index_boundary is not None and node.index_boundary not in seen:
                    for i in range(6):
                        if not node.fixed_DOF[i]:
                            out.append(node.displacement_vector[i])
                            idx.append(node.index_boundary)
                    seen.add(node.index_boundary)
        self.global_displacement_index = idx
        return np.array(out), idx


# ---------------------------------------------------------------------------------------------------------
# Reference-shaped objects rebuilt FROM A DUMP of the reference object graph (tests/golden/objgraph_*.npz,
# written by tests/golden/make_golden.py::make_objgraph from real pyLatticeDSO objects).  Nothing here comes
# from the repo's own generator: coordinates, numbering, beam radii / penalisation flags, cell membership,
# index_boundary, fixed_DOF, imposed values, loads and the (node, DOF) order of get_global_displacement are
# all read from the dump.
# ---------------------------------------------------------------------------------------------------------
class DumpCell:
    def __init__(self, index, center, radii):
        self.index = int(index)
        self.center_point = tuple(float(v) for v in center)
        self.radii = [float(v) for v in radii]
        self.points_cell, self.beams_cell = set(), set()
        self.node_in_order_simulation = None
        self.schur_complement = None
        self.cell_size = (1.0, 1.0, 1.0)

    def define_node_order_to_simulate(self):   # cell.py:611-680 (face priority, in-plane sort) on the dumped nodes
        pts = [p for p in self.points_cell if p.index_boundary is not None]
        xyz = np.array([[p.x, p.y, p.z] for p in pts])
        h = [0.5 * s for s in self.cell_size]
        c = self.center_point
        box = (c[0] - h[0], c[0] + h[0], c[1] - h[1], c[1] + h[1], c[2] - h[2], c[2] + h[2])
        order = bcc_cell_order_nodes(xyz, box)
        self.node_in_order_simulation = [pts[i] for i in order]


class DumpLattice:
    material_name = "VeroClear"

    def __init__(self, G):
        self.cell_size_x = float(G["cell_size_x"])
        self.max_index_boundary = int(G["max_index_boundary"])
        self.geom_types = list(range(int(G["n_geom"])))
        self.points = {}
        for k, idx in enumerate(G["p_index"]):
            p = FakePoint(*G["p_xyz"][k], idx)
            ib = int(G["p_index_boundary"][k])
            p.index_boundary = None if ib < 0 else ib
            p.fixed_DOF = [int(v) for v in G["p_fixed"][k]]
            p.displacement_vector = [float(v) for v in G["p_imposed"][k]]
            p.applied_force = [float(v) for v in G["p_force"][k]]
            self.points[int(idx)] = p
        self.beams = {}
        for k, idx in enumerate(G["b_index"]):
            b = FakeBeam(self.points[int(G["b_p1"][k])], self.points[int(G["b_p2"][k])], G["b_radius"][k], idx,
                         type_beam=int(G["b_type"][k]))
            b.beam_mod = bool(G["b_mod"][k])
            b.penalization_coefficient = float(G["b_pen"][k])
            self.beams[int(idx)] = b
        self.cells = []
        for k, idx in enumerate(G["c_index"]):
            c = DumpCell(idx, G["c_center"][k], G["c_radii"][k])
            c.points_cell = {self.points[int(i)] for i in G["c_points"][G["c_points_ptr"][k]: G["c_points_ptr"][k + 1]]}
            c.beams_cell = {self.beams[int(i)] for i in G["c_beams"][G["c_beams_ptr"][k]: G["c_beams_ptr"][k + 1]]}
            self.cells.append(c)
        self._order = [(int(n), int(d)) for n, d in G["xsol_order"]]
        self.global_displacement_index = None

    def get_number_cells(self):
        return len(self.cells)

    def get_global_displacement(self):
        """LatticeSim.get_global_displacement (lattice_sim.py:502-542) in the ORDER the reference produced (dumped)."""
        vals = [self.points[n].displacement_vector[d] for n, d in self._order]
        self.global_displacement_index = [self.points[n].index_boundary for n, _ in self._order]
        return np.array(vals), self.global_displacement_index


def lattice_from_dump(G):
    return DumpLattice(G)


class DumpDdmLattice:
    """DDM view of a reference lattice rebuilt from tests/golden/objgraph_ddm_*.npz: cell-boundary nodes, the
    reference's free-DOF numbering (Point.global_free_DOF_index), per-cell boundary-node order and Schur matrix."""
    material_name = "VeroClear"

    def __init__(self, G):
        self.cell_size_x = float(G["cell_size_x"])
        self.max_index_boundary = int(G["max_index_boundary"])
        self._free_DOF = int(G["free_DOF"])
        self.free_DOF = None
        self.points = {}
        self._gfree = {}
        for k, idx in enumerate(G["p_index"]):
            p = FakePoint(*G["p_xyz"][k], idx)
            ib = int(G["p_index_boundary"][k])
            p.index_boundary = None if ib < 0 else ib
            p.fixed_DOF = [int(v) for v in G["p_fixed"][k]]
            p.displacement_vector = [float(v) for v in G["p_imposed"][k]]
            p.applied_force = [float(v) for v in G["p_force"][k]]
            self.points[int(idx)] = p
            self._gfree[int(idx)] = [None if v < 0 else int(v) for v in G["p_free_index"][k]]
        S = np.array(G["schur_shared"])
        self.cells = []
        for k, idx in enumerate(G["c_index"]):
            c = DumpCell(idx, (0.0, 0.0, 0.0), [0.05])
            c.node_in_order_simulation = [self.points[int(i)] for i in G["cell_node_order"][k]]
            c.points_cell = set(c.node_in_order_simulation)
            c.schur_complement = S
            c.define_node_order_to_simulate = lambda: None
            self.cells.append(c)
        self._order = [(int(n), int(d)) for n, d in G["xsol_order"]]
        self.global_displacement_index = None
        self.python_loop_calls = 0

    def define_free_DOF(self):
        self.free_DOF = self._free_DOF

    def set_global_free_DOF_index(self):
        for idx, p in self.points.items():
            p.global_free_DOF_index = list(self._gfree[idx])

    def get_global_displacement(self):
        vals = [self.points[n].displacement_vector[d] for n, d in self._order]
        self.global_displacement_index = [self.points[n].index_boundary for n, _ in self._order]
        return np.array(vals), self.global_displacement_index

    def calculate_reaction_force_global(self, v, rightHandSide=False):
        """The reference's Python loop over all cells (lattice_sim.py:1180-1252): must NEVER run on the B200 path."""
        self.python_loop_calls += 1
        raise AssertionError("the device path called the Python cell loop")
