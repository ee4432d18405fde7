"""Generate the committed fixtures under tests/golden/ (run HERE, where a
pyLatticeDSO checkout exists at /root/reference; the GPU box has none).

    python tests/golden/make_golden.py

What it freezes:

* schur_<geom>.npz       the reference's 30 stored dolfinx/PETSc Schur matrices
                         (data/outputs/schur_complement/*.npz) together with the
                         flattened 1-cell meshes (reference object graph ->
                         arrays) and the boundary-DOF order they were computed
                         for.  These PIN the oracle.
* numbering_*.npz        node/beam numbering of small reference lattices
                         (pyLatticeDesign object graph), to check the vectorised
                         generator without the reference.
* ddm_loop_*.npz         reference-in-the-loop displacements: the reference's OWN
                         LatticeSim.solve_DDM (lattice_sim.py:1111-1176) run with
                         get_schur_complement rebound to the oracle's cell Schur
                         (which the 30 goldens pin), on a 3x2x2 penalised BCC
                         lattice with the reference gmsh subdivision.
* grad_loop_*.npz        reference-in-the-loop compliance and gradient: the
                         reference's OWN LatticeOpti.objective/gradient
                         (lattice_opti.py:430-465,701-731) on a 3x1x1 BCC lattice.
* objgraph_*.npz         the reference OBJECT GRAPH of a penalised 3x2x2 BCC lattice and of a graded
                         2x2x3 Octet lattice (points, beams, cells, index_boundary, fixed_DOF, imposed
                         values, loads) plus what the reference's own write-back leaves on it when
                         fed the oracle's FEM solution: Point.displacement_vector, the k-fold
                         accumulated Point.reaction_force_vector (point.py:372-385 through
                         full_scale_lattice_simulation.py:111-120) and LatticeSim.get_global_displacement()
                         with its (node, DOF) order.  The GPU drop-in test rebuilds duck-typed objects
                         FROM THIS DUMP (tests/fake_lattice.lattice_from_dump), not from the repo's generator.
* pcg_reference.npz      inputs/outputs of the reference's OWN
                         conjugate_gradient_solver (conjugate_gradient_solver.py)
                         on small SPD systems, including alpha-clamp and restart.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from pylatticedso_b200 import refshim  # noqa: E402
from pylatticedso_b200 import mesh as M  # noqa: E402
from oracle import lattice_oracle as orc  # noqa: E402

E_MOD, NU = 1013.0, 0.3  # materials/VeroClear.json:3-5


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def base_cfg(geom, n, radii, enable, periodicity, extra=None):
    geoms = geom if isinstance(geom, list) else [geom]
    cfg = {"geometry": {"cell_size": {"x": 1, "y": 1, "z": 1},
                        "number_of_cells": {"x": n[0], "y": n[1], "z": n[2]},
                        "radii": list(radii), "geom_types": geoms},
           "simulation_parameters": {"enable": enable, "material": "VeroClear", "periodicity": periodicity}}
    if extra:
        for k, v in extra.items():
            if k == "simulation_parameters":
                cfg[k].update(v)
            else:
                cfg[k] = v
    return cfg


def mesh_arrays(mesh, prefix=""):
    return {prefix + "x": mesh.x, prefix + "y": mesh.y, prefix + "z": mesh.z,
            prefix + "en0": mesh.en0, prefix + "en1": mesh.en1, prefix + "rad": mesh.rad,
            prefix + "chain": mesh.chain, prefix + "beam_of_elem": mesh.beam_of_elem,
            prefix + "cell_of_elem": mesh.cell_of_elem, prefix + "n_points": np.int64(mesh.n_points),
            prefix + "point_index": mesh.point_index}


def make_schur(ls):
    for geom, enable in (("BCC", True), ("Hybrid1", False), ("Hybrid4", False)):
        G = np.load(os.path.join(refshim.reference_root(), "data", "outputs", "schur_complement",
                                 f"Schur_complement_{geom}.npz"))
        rv, SM = G["radius_values"], G["schur_matrices"]
        out = {"radius_values": rv, "schur_matrices": SM, "enable_penalization": np.bool_(enable)}
        worst = 0.0
        for i in range(len(rv)):
            refshim.set_inline_presets({"g": base_cfg(geom, (1, 1, 1), [float(rv[i][0])], enable, True)})
            with quiet():
                lat = ls.LatticeSim("g")
            mesh = M.flatten_lattice(lat, 0, "gmsh")
            bnd = M.cell_boundary_dofs(lat.cells[0], mesh)
            out.update(mesh_arrays(mesh, f"c{i}_"))
            out[f"c{i}_bnd"] = bnd
            K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
            S = orc.schur_complement(K, bnd)
            worst = max(worst, np.abs(S - SM[i]).max() / np.abs(SM[i]).max())
        print(f"schur_{geom}: oracle vs golden worst rel err {worst:.2e}")
        np.savez_compressed(os.path.join(HERE, f"schur_{geom}.npz"), **out)


def make_numbering(ls):
    cases = {
        "bcc_322": ("BCC", (3, 2, 2), [0.05], None),
        "octet_322": ("Octet", (3, 2, 2), [0.03], None),
        "octet_223_graded": ("Octet", (2, 2, 3), [0.03],
                             {"radii": {"rule": "linear", "direction_z": True, "parameter_z": 0.0125}}),
        "bcc_345": ("BCC", (3, 4, 5), [0.04], None),
    }
    for name, (geom, n, radii, grad) in cases.items():
        extra = {"gradient": grad} if grad else None
        refshim.set_inline_presets({"s": base_cfg(geom, n, radii, False, False, extra)})
        with quiet():
            lat = ls.LatticeSim("s")
        nodes = sorted(lat.nodes, key=lambda p: p.index)
        beams = sorted(lat.beams, key=lambda b: b.index)
        cell_radii = np.array([[c.get_radius(r) for r in c.radii] for c in lat.cells])
        np.savez_compressed(
            os.path.join(HERE, f"numbering_{name}.npz"),
            geom=geom, n_cells=np.array(n), radii=np.array(radii),
            grad_param_z=np.float64(grad["radii"]["parameter_z"] if grad else 0.0),
            pxyz=np.array([[p.x, p.y, p.z] for p in nodes]),
            b_p1=np.array([b.point1.index for b in beams]), b_p2=np.array([b.point2.index for b in beams]),
            b_rad=np.array([b.radius for b in beams]),
            b_cell=np.array([min(c.index for c in b.cell_belongings) for b in beams]),
            cell_radii=cell_radii)
        print(f"numbering_{name}: {len(nodes)} nodes {len(beams)} beams")


def _oracle_schur_rebind(ls, elements_per_strut="gmsh"):
    def _schur(lattice, cell_index=None):
        idx = 0 if cell_index is None else cell_index
        return orc.cell_schur_from_lattice(lattice, idx, E_MOD, NU, elements_per_strut)
    ls.get_schur_complement = _schur


def make_ddm_loop(ls):
    ddm = {"DDM": {"enable_preconditioner": True, "preconditioner_type": "exact", "max_iterations": 500,
                   "schur_complement_computation": {"type": "exact"}}}
    cases = {
        "disp": {"Displacement": {
            "Fixed": {"Surface": ["Xmin"], "DOF": ["X", "Y", "Z", "RX", "RY", "RZ"], "Value": [0, 0, 0, 0, 0, 0]},
            "Load": {"Surface": ["Xmax"], "DOF": ["Z"], "Value": [-0.01]}}},
        "force": {"Displacement": {
            "Fixed": {"Surface": ["Zmin"], "DOF": ["X", "Y", "Z", "RX", "RY", "RZ"], "Value": [0, 0, 0, 0, 0, 0]}},
            "Force": {"Load": {"Surface": ["Zmax"], "DOF": ["Z", "X"], "Value": [-0.5, 0.2]}}},
    }
    orig = ls.get_schur_complement
    try:
        _oracle_schur_rebind(ls)
        for name, bc in cases.items():
            cfg = base_cfg("BCC", (3, 2, 2), [0.05], True, True,
                           {"simulation_parameters": ddm, "boundary_conditions": bc})
            refshim.set_inline_presets({"d": cfg})
            with quiet():
                lat = ls.LatticeSim("d", enable_domain_decomposition_solver=True)
                xsol, info, _, b = lat.solve_DDM()
            mesh = M.flatten_lattice(lat, None, "gmsh")
            fixed, g, f = M.bc_arrays_from_lattice(lat, mesh, dedup_point_loads=True)
            fixed_q, g_q, f_q = M.bc_arrays_from_lattice(lat, mesh, dedup_point_loads=False)
            pts = mesh.meta["points"]
            u_pts = np.array([p.displacement_vector for p in pts], dtype=np.float64)
            has_bnd = np.array([p.index_boundary is not None for p in pts])
            # independent check with the oracle's full FEM
            K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
            u, R = orc.solve_static(K, fixed.astype(bool), g, f)
            uo = u.reshape(-1, 6)[: mesh.n_points]
            err = np.abs(uo[has_bnd] - u_pts[has_bnd]).max() / np.abs(u_pts).max()
            print(f"ddm_loop_{name}: info={info} n_dof={mesh.n_dof} oracle-FEM vs reference-DDM rel err {err:.2e}")
            np.savez_compressed(os.path.join(HERE, f"ddm_loop_{name}.npz"), fixed=fixed, g=g, f=f,
                                f_reference_fem_quirk=f_q, u_points_reference_ddm=u_pts,
                                point_on_cell_boundary=has_bnd, xsol=np.asarray(xsol), info=np.int64(info),
                                **mesh_arrays(mesh))
    finally:
        ls.get_schur_complement = orig


def make_grad_loop(ls):
    import importlib
    lo = importlib.import_module("pyLatticeOpti.lattice_opti")
    refshim.set_inline_presets({})  # make sure lattice_opti's open_lattice_parameters is routed too
    ddm = {"DDM": {"enable_preconditioner": True, "preconditioner_type": "exact", "max_iterations": 500,
                   "schur_complement_computation": {"type": "exact"}}}
    bc = {"Displacement": {"Fixed": {"Surface": ["Xmin"], "DOF": ["X", "Y", "Z", "RX", "RY", "RZ"],
                                     "Value": [0, 0, 0, 0, 0, 0]}},
          "Force": {"Load": {"Surface": ["Xmax"], "DOF": ["Z"], "Value": [-0.1]}}}
    opt = {"objective_function": "min", "objective_type": "compliance", "max_iterations": 2,
           "optimization_parameters": {"type": "unit_cell", "hybrid": False},
           "enable_parameter_normalization": False, "simulation_type": "DDM",
           "enable_gradient_computing": True}
    cfg = base_cfg("BCC", (3, 1, 1), [0.05], True, True,
                   {"simulation_parameters": ddm, "boundary_conditions": bc, "optimization_informations": opt})
    refshim.set_inline_presets({"o": cfg})
    orig = ls.get_schur_complement
    try:
        _oracle_schur_rebind(ls)
        r = [0.04, 0.05, 0.06]
        with quiet():
            lat = lo.LatticeOpti("o")
            lat.enable_normalization = False
            obj = lat.objective(np.array(r))
            compliance = lat.compute_compliance()
            grad = np.asarray(lat.gradient(np.array(r)), dtype=np.float64)
        mesh = M.flatten_lattice(lat, None, "gmsh")
        fixed, g, f = M.bc_arrays_from_lattice(lat, mesh, dedup_point_loads=True)
        # group of each element = cell that owns its beam (BCC: struts are private to a cell)
        group = mesh.cell_of_elem.copy()
        # analytic oracle for comparison
        K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
        u, _ = orc.solve_static(K, fixed.astype(bool), g, f)
        ga = orc.compliance_gradient(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, u, group, 3,
                                     E_MOD, NU, chain=mesh.chain)
        print("grad_loop: reference compliance", compliance, "oracle", float(f @ u))
        print("grad_loop: reference FD gradient", grad, "oracle analytic", ga,
              "max |diff| / |g|_inf", np.abs(grad - ga).max() / np.abs(grad).max())
        np.savez_compressed(os.path.join(HERE, "grad_loop_bcc311.npz"), fixed=fixed, g=g, f=f,
                            cell_radii=np.array(r), compliance_reference=np.float64(compliance),
                            objective_reference=np.float64(obj), gradient_reference_fd=grad,
                            gradient_oracle_analytic=ga, group=group, **mesh_arrays(mesh))
    finally:
        ls.get_schur_complement = orig


def make_c1_parity(ls):
    """BASELINE configs[0] in parity mode: BCC 5x5x5, r = 0.05, joint penalisation on, reference gmsh
    subdivision (h = 0.05): 2 992 beams -> 18 992 elements / 18 333 nodes / 109 998 DOF, uniaxial compression.
    Mesh and BCs come from the reference object graph; the stored solution is the oracle's direct solve."""
    bc = {"Displacement": {
        "Fixed": {"Surface": ["Zmin"], "DOF": ["X", "Y", "Z", "RX", "RY", "RZ"], "Value": [0, 0, 0, 0, 0, 0]},
        "Load": {"Surface": ["Zmax"], "DOF": ["Z"], "Value": [-0.01]}}}
    refshim.set_inline_presets({"c1": base_cfg("BCC", (5, 5, 5), [0.05], True, False, {"boundary_conditions": bc})})
    with quiet():
        lat = ls.LatticeSim("c1")
    mesh = M.flatten_lattice(lat, None, "gmsh")
    fixed, g, f = M.bc_arrays_from_lattice(lat, mesh)
    K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
    u, R = orc.solve_static(K, fixed.astype(bool), g, f)
    npnt = mesh.n_points
    print(f"c1_parity: beams={len(lat.beams)} nodes={mesh.n_nodes} elements={mesh.n_elems} dof={mesh.n_dof} "
          f"fixed={int(fixed.sum())} |u|max={np.abs(u).max():.4e}")
    arr = mesh_arrays(mesh)
    arr["rad"] = arr["rad"].astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "c1_parity_bcc555.npz"), fixed=np.packbits(fixed), g_nonzero_idx=np.flatnonzero(g),
                        g_nonzero_val=g[np.flatnonzero(g)], u_points_oracle=u.reshape(-1, 6)[:npnt],
                        reactions_points_oracle=R.reshape(-1, 6)[:npnt], n_dof=np.int64(mesh.n_dof), **arr)


def make_objgraph(ls):
    """Dump the reference object graph + the result of the reference's own write-back methods (module docstring)."""
    cases = {
        "bcc322_pen": (base_cfg("BCC", (3, 2, 2), [0.05], True, True, {"boundary_conditions": {
            "Displacement": {"Fixed": {"Surface": ["Zmin"], "DOF": ["X", "Y", "Z", "RX", "RY", "RZ"], "Value": [0, 0, 0, 0, 0, 0]},
                             "Load": {"Surface": ["Zmax"], "DOF": ["Z"], "Value": [-0.01]}},
            "Force": {"Push": {"Surface": ["Xmax"], "DOF": ["X", "Y"], "Value": [0.3, -0.1]}}}}), {}),
        "octet223_graded": (base_cfg("Octet", (2, 2, 3), [0.03], False, False, {
            "gradient": {"radii": {"rule": "linear", "direction": [False, False, True], "parameters": [0.0, 0.0, 0.2]}},
            "boundary_conditions": {
                "Displacement": {"Fixed": {"Surface": ["Xmin"], "DOF": ["X", "Y", "Z", "RX", "RY", "RZ"], "Value": [0, 0, 0, 0, 0, 0]}},
                "Force": {"Pull": {"Surface": ["Xmax"], "DOF": ["Z"], "Value": [-0.2]}}}}), {}),
    }
    for name, (cfg, _) in cases.items():
        refshim.set_inline_presets({"g": cfg})
        with quiet():
            lat = ls.LatticeSim("g", enable_domain_decomposition_solver=False)
        mesh = M.flatten_lattice(lat, None, "gmsh")
        pts = mesh.meta["points"]                       # node.index order
        loc = {p.index: k for k, p in enumerate(pts)}
        beams = {}
        for c in lat.cells:
            for b in c.beams_cell:
                beams[b.index] = b
        border = sorted(beams)
        # ---- inputs
        P = dict(p_index=np.array([p.index for p in pts], dtype=np.int64),
                 p_xyz=np.array([[p.x, p.y, p.z] for p in pts], dtype=np.float64),
                 p_index_boundary=np.array([-1 if p.index_boundary is None else p.index_boundary for p in pts], dtype=np.int64),
                 p_fixed=np.array([[int(bool(v)) for v in p.fixed_DOF] for p in pts], dtype=np.uint8),
                 p_imposed=np.array([list(p.displacement_vector) for p in pts], dtype=np.float64),
                 p_force=np.array([list(p.applied_force) for p in pts], dtype=np.float64))
        B = dict(b_index=np.array(border, dtype=np.int64),
                 b_p1=np.array([beams[i].point1.index for i in border], dtype=np.int64),
                 b_p2=np.array([beams[i].point2.index for i in border], dtype=np.int64),
                 b_radius=np.array([beams[i].radius for i in border], dtype=np.float64),
                 b_mod=np.array([bool(getattr(beams[i], "beam_mod", False)) for i in border], dtype=np.uint8),
                 b_type=np.array([int(getattr(beams[i], "type_beam", 0)) for i in border], dtype=np.int64),
                 b_pen=np.array([float(getattr(beams[i], "penalization_coefficient", 1.5)) for i in border], dtype=np.float64))
        cp_ptr, cp, cb_ptr, cb = [0], [], [0], []
        for c in lat.cells:
            cp.extend(sorted(p.index for p in c.points_cell)); cp_ptr.append(len(cp))
            cb.extend(sorted(b.index for b in c.beams_cell)); cb_ptr.append(len(cb))
        Cc = dict(c_index=np.array([c.index for c in lat.cells], dtype=np.int64),
                  c_points_ptr=np.array(cp_ptr, dtype=np.int64), c_points=np.array(cp, dtype=np.int64),
                  c_beams_ptr=np.array(cb_ptr, dtype=np.int64), c_beams=np.array(cb, dtype=np.int64),
                  c_center=np.array([list(c.center_point) for c in lat.cells], dtype=np.float64),
                  c_radii=np.array([list(c.radii) for c in lat.cells], dtype=np.float64))
        # ---- (node, DOF) order of get_global_displacement: inject a code into every displacement entry
        saved = [list(p.displacement_vector) for p in pts]
        for p in pts:
            p.displacement_vector[:] = [float(p.index * 6 + d + 1) for d in range(6)]
        with quiet():
            codes, gdi = lat.get_global_displacement()
        order = np.array([[(int(c) - 1) // 6, (int(c) - 1) % 6] for c in codes], dtype=np.int64)
        for p, v in zip(pts, saved):
            p.displacement_vector[:] = v
        # ---- the oracle's FEM solution, written back by the reference's own objects
        fixed, g, f = M.bc_arrays_from_lattice(lat, mesh)            # reference quirk kept: loads once per owning cell
        K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
        u, R = orc.solve_static(K, fixed.astype(bool), g, f)
        un, Rn = u.reshape(-1, 6), R.reshape(-1, 6)
        for k, p in enumerate(pts):
            p.displacement_vector[:] = [float(v) for v in un[k]]
            p.reaction_force_vector = [0.0] * 6
        for c in lat.cells:                                          # full_scale_lattice_simulation.py:111-120
            for node in c.points_cell:
                if 1 in node.fixed_DOF:
                    node.set_reaction_force([float(v) for v in Rn[loc[node.index]]])
        with quiet():
            xsol, gdi = lat.get_global_displacement()
        out = dict(P, **B, **Cc, xsol_order=order, xsol_expected=np.asarray(xsol, dtype=np.float64),
                   global_displacement_index=np.asarray(gdi, dtype=np.int64),
                   u_points_expected=np.array([list(p.displacement_vector) for p in pts], dtype=np.float64),
                   reaction_points_expected=np.array([list(p.reaction_force_vector) for p in pts], dtype=np.float64),
                   cell_size_x=np.float64(lat.cell_size_x), max_index_boundary=np.int64(lat.max_index_boundary),
                   n_geom=np.int64(len(lat.geom_types)))
        np.savez_compressed(os.path.join(HERE, f"objgraph_{name}.npz"), **out)
        kmax = max(sum(1 for c in lat.cells if p in c.points_cell) for p in pts)
        print(f"objgraph_{name}: points={len(pts)} beams={len(border)} cells={len(lat.cells)} fe_dof={mesh.n_dof} "
              f"fixed={int(fixed.sum())} xsol={len(xsol)} max cells per node={kmax}")


def make_ddm_objects(ls):
    """objgraph_ddm_bcc322.npz: the DDM view of a penalised 3x2x2 BCC lattice (periodic joints: one Schur matrix shared
    by all cells, computed by the oracle) together with what the REFERENCE'S OWN code computes on it:
    y = LatticeSim.calculate_reaction_force_global(v) for a random v (lattice_sim.py:1180-1252, the Python loop over
    cells), and (xsol, b) of LatticeSim.solve_DDM (:1111-1176, exact preconditioner)."""
    ddm = {"DDM": {"enable_preconditioner": True, "preconditioner_type": "exact", "max_iterations": 500,
                   "schur_complement_computation": {"type": "exact"}}}
    bc = {"Displacement": {"Fixed": {"Surface": ["Xmin"], "DOF": ["X", "Y", "Z", "RX", "RY", "RZ"], "Value": [0, 0, 0, 0, 0, 0]},
                           "Load": {"Surface": ["Xmax"], "DOF": ["Z"], "Value": [-0.01]}},
          "Force": {"Push": {"Surface": ["Zmax"], "DOF": ["Y"], "Value": [0.05]}}}
    cfg = base_cfg("BCC", (3, 2, 2), [0.05], True, True, {"simulation_parameters": ddm, "boundary_conditions": bc})
    refshim.set_inline_presets({"d": cfg})
    orig = ls.get_schur_complement
    try:
        _oracle_schur_rebind(ls)
        with quiet():
            lat = ls.LatticeSim("d", enable_domain_decomposition_solver=True)
            xsol, info, gdi, b = lat.solve_DDM()
        u_after = {}
        for c in lat.cells:
            for p in c.points_cell:
                if p.index_boundary is not None:
                    u_after[p.index] = (list(p.displacement_vector), list(p.reaction_force_vector))
        rng = np.random.default_rng(11)
        v = rng.standard_normal(lat.free_DOF) * 1e-3
        with quiet():
            lat._initialize_displacement()
            y = np.asarray(lat.calculate_reaction_force_global(v), dtype=np.float64)
            lat._initialize_displacement()
            lat.set_boundary_conditions()
        mesh = M.flatten_lattice(lat, None, "gmsh")
        pts = mesh.meta["points"]
        S = np.asarray(lat.cells[0].schur_complement, dtype=np.float64)
        assert all(np.array_equal(np.asarray(c.schur_complement), S) for c in lat.cells)
        order = np.array([[p.index for p in c.node_in_order_simulation] for c in lat.cells], dtype=np.int64)
        gfree = np.array([[-1 if (p.fixed_DOF[d] or p.index_boundary is None or p.global_free_DOF_index[d] is None)
                           else int(p.global_free_DOF_index[d]) for d in range(6)] for p in pts], dtype=np.int64)
        out = dict(p_index=np.array([p.index for p in pts], dtype=np.int64),
                   p_xyz=np.array([[p.x, p.y, p.z] for p in pts]),
                   p_index_boundary=np.array([-1 if p.index_boundary is None else p.index_boundary for p in pts], dtype=np.int64),
                   p_fixed=np.array([[int(bool(q)) for q in p.fixed_DOF] for p in pts], dtype=np.uint8),
                   p_imposed=np.array([list(p.displacement_vector) for p in pts]),
                   p_force=np.array([list(p.applied_force) for p in pts]),
                   p_free_index=gfree, cell_node_order=order, c_index=np.array([c.index for c in lat.cells], dtype=np.int64),
                   schur_shared=S, v=v, y_reference=y, xsol_reference=np.asarray(xsol, dtype=np.float64),
                   b_reference=np.asarray(b, dtype=np.float64), info_reference=np.int64(info),
                   global_displacement_index=np.asarray(gdi, dtype=np.int64), free_DOF=np.int64(lat.free_DOF),
                   max_index_boundary=np.int64(lat.max_index_boundary), cell_size_x=np.float64(lat.cell_size_x))
        # (node, DOF) order of get_global_displacement, as in make_objgraph
        saved = [list(p.displacement_vector) for p in pts]
        for p in pts:
            p.displacement_vector[:] = [float(p.index * 6 + d + 1) for d in range(6)]
        with quiet():
            codes, _ = lat.get_global_displacement()
        out["xsol_order"] = np.array([[(int(c) - 1) // 6, (int(c) - 1) % 6] for c in codes], dtype=np.int64)
        for p, q in zip(pts, saved):
            p.displacement_vector[:] = q
        bnd = [p for p in pts if p.index_boundary is not None]
        out["u_boundary_reference"] = np.array([u_after[p.index][0] for p in bnd])
        out["boundary_point_index"] = np.array([p.index for p in bnd], dtype=np.int64)
        np.savez_compressed(os.path.join(HERE, "objgraph_ddm_bcc322.npz"), **out)
        print(f"objgraph_ddm_bcc322: free_DOF={lat.free_DOF} cells={len(lat.cells)} info={info} |y|={np.abs(y).max():.3e} "
              f"|xsol|={np.abs(xsol).max():.3e}")
    finally:
        ls.get_schur_complement = orig


def make_pcg(ls):
    import importlib
    cgm = importlib.import_module("pyLatticeSim.conjugate_gradient_solver")
    rng = np.random.default_rng(7)
    n = 60
    Q = rng.standard_normal((n, n))
    A = Q @ Q.T + n * np.eye(n)
    A2 = Q @ np.diag(np.logspace(0, 4, n)) @ Q.T
    A2 = (A2 + A2.T) / 2 + 1e-3 * np.eye(n)
    out = {"A_well": A, "A_ill": A2}
    cases = [
        ("well_default", A, dict(maxiter=100, tol=1e-5, mintol=1e-5, restart_every=1000, alpha_max=0.1), True),
        ("well_ddm", A, dict(maxiter=200, tol=1e-6, mintol=1e-12, restart_every=500000, alpha_max=100), True),
        ("ill_clamp", A2, dict(maxiter=40, tol=1e-10, mintol=1e-14, restart_every=7, alpha_max=0.05), False),
        ("ill_restart", A2, dict(maxiter=300, tol=1e-8, mintol=1e-14, restart_every=5, alpha_max=100), True),
    ]
    for name, mat, kw, jac in cases:
        b = rng.standard_normal(n)
        Minv = np.diag(1.0 / np.diag(mat)) if jac else None
        with quiet():
            x, info = cgm.conjugate_gradient_solver(mat, b, M=Minv, **kw)
        xo, io_, it = orc.reference_pcg(mat, b, Minv, **kw)
        print(f"pcg {name}: info={info} oracle info={io_} iters={it} |x-xo|={np.abs(x - xo).max():.1e}")
        out[f"{name}_b"] = b
        out[f"{name}_x"] = x
        out[f"{name}_info"] = np.int64(info)
        out[f"{name}_iters"] = np.int64(it)
        out[f"{name}_jacobi"] = np.bool_(jac)
        out[f"{name}_params"] = np.array([kw["maxiter"], kw["tol"], kw["mintol"], kw["restart_every"], kw["alpha_max"]])
    np.savez_compressed(os.path.join(HERE, "pcg_reference.npz"), **out)


def main():
    ls = refshim.import_reference()
    make_schur(ls)
    make_numbering(ls)
    make_ddm_loop(ls)
    make_grad_loop(ls)
    make_c1_parity(ls)
    make_objgraph(ls)
    make_ddm_objects(ls)
    make_pcg(ls)
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print(f"total fixture size {tot / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
