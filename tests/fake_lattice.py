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
