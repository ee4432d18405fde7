"""CPU, only where a pyLatticeDSO checkout exists (/root/reference in the build container; the GPU box has none):
install.patch_reference() rebinds EVERY seam of SURVEY section 8(b) on the real reference modules, the rebound
names reach the B200 layer (which fails loudly without a GPU instead of falling back), and the host-side parameter
mapping reproduces the reference's own calculate_gradient loops on real LatticeOpti objects."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "pyLatticeSim")),
                                reason="no pyLatticeDSO checkout on this machine")

EXPECTED = {
    "pyLatticeSim.utils_simulation.solve_FEM_FenicsX", "pyLatticeOpti.lattice_opti.solve_FEM_FenicsX",
    "pyLatticeSim.utils_schur.get_schur_complement", "pyLatticeSim.lattice_sim.get_schur_complement",
    "pyLatticeSim.conjugate_gradient_solver.conjugate_gradient_solver", "pyLatticeSim.lattice_sim.conjugate_gradient_solver",
    "pyLatticeOpti.lattice_opti.conjugate_gradient_solver",
    "pyLatticeSim.lattice_sim.LatticeSim.solve_DDM", "pyLatticeSim.lattice_sim.LatticeSim._compute_schur_gradients",
    "pyLatticeOpti.lattice_opti.LatticeOpti.calculate_gradient",
    # N4: reduced-basis / surrogate pipeline
    "pyLatticeSim.greedy_algorithm.reduce_basis_greedy", "pyLatticeSim.greedy_algorithm.project_to_reduced_basis",
    "pyLatticeSim.utils_rbf.ThinPlateSplineRBF", "pyLatticeSim.lattice_sim.ThinPlateSplineRBF",
    "pyLatticeSim.lattice_sim.LatticeSim.get_schur_complement_from_reduced_basis_batch",
    "pyLatticeSim.lattice_sim.LatticeSim.get_schur_complement_from_reduced_basis",
    "pyLatticeSim.lattice_sim.LatticeSim._compute_schur_gradients_RBF",
    "pyLatticeSim.lattice_sim.LatticeSim._define_radial_basis_functions",
}


@pytest.fixture(scope="module")
def ref():
    from pylatticedso_b200 import refshim
    ls = refshim.import_reference()
    import importlib
    lo = importlib.import_module("pyLatticeOpti.lattice_opti")
    importlib.import_module("pyLatticeSim.utils_simulation")
    importlib.import_module("pyLatticeSim.utils_schur")
    refshim.set_inline_presets({})
    return refshim, ls, lo


def _cfg(n, opt=None, ddm=True):
    cfg = {"geometry": {"cell_size": {"x": 1, "y": 1, "z": 1}, "number_of_cells": {"x": n[0], "y": n[1], "z": n[2]},
                        "radii": [0.05], "geom_types": ["BCC"]},
           "simulation_parameters": {"enable": True, "material": "VeroClear", "periodicity": True},
           "boundary_conditions": {
               "Displacement": {"Fixed": {"Surface": ["Xmin"], "DOF": ["X", "Y", "Z", "RX", "RY", "RZ"], "Value": [0] * 6}},
               "Force": {"Load": {"Surface": ["Xmax"], "DOF": ["Z"], "Value": [-0.1]}}}}
    if ddm:
        cfg["simulation_parameters"]["DDM"] = {"enable_preconditioner": True, "preconditioner_type": "exact", "max_iterations": 500,
                                               "schur_complement_computation": {"type": "exact"}}
    if opt:
        cfg["optimization_informations"] = opt
    return cfg


def test_patch_reference_rebinds_every_seam_and_fails_loudly_without_gpu(ref):
    import torch
    from pylatticedso_b200 import install
    from pylatticedso_b200.lib import LatticeB200Error
    refshim, ls, lo = ref
    before = {n: getattr(sys.modules[n.rsplit(".", 1)[0]], n.rsplit(".", 1)[1]) for n in EXPECTED if n.count(".") == 2}
    done = install.patch_reference()
    try:
        assert set(done) == EXPECTED
        assert ls.LatticeSim.solve_DDM.__name__ == "_solve_ddm"
        assert lo.LatticeOpti.calculate_gradient.__name__ == "_calculate_gradient"
        if not torch.cuda.is_available():
            refshim.set_inline_presets({"t": _cfg((2, 1, 1), ddm=False)})
            lat = ls.LatticeSim("t")
            with pytest.raises(LatticeB200Error):            # the B200 layer, not the reference's dolfinx path
                sys.modules["pyLatticeSim.utils_simulation"].solve_FEM_FenicsX(lat)
            with pytest.raises(LatticeB200Error):
                ls.get_schur_complement(lat, 0)
            with pytest.raises(LatticeB200Error):            # surrogate seam: the device layer, not scipy BLAS
                sys.modules["pyLatticeSim.greedy_algorithm"].reduce_basis_greedy({(0.1,): np.eye(6), (0.2,): 2 * np.eye(6)}, 1e-3)
            with pytest.raises(LatticeB200Error):
                ls.ThinPlateSplineRBF(np.array([[0.0], [1.0], [2.0]]), np.array([1.0, 2.0, 4.0]))
    finally:
        assert install.unpatch_reference() == len(EXPECTED)
    for n, fn in before.items():
        assert getattr(sys.modules[n.rsplit(".", 1)[0]], n.rsplit(".", 1)[1]) is fn


def test_operator_recognition_on_the_reference_linear_operators(ref):
    """The two ways the reference wraps calculate_reaction_force_global (lattice_sim.py:1148, lattice_opti.py:1640)."""
    from scipy.sparse.linalg import LinearOperator
    from pylatticedso_b200.pcg import lattice_of_operator
    refshim, ls, lo = ref

    class Dummy:
        cells = []
        def calculate_reaction_force_global(self, v, rightHandSide=False): return v
    d = Dummy()
    A = LinearOperator(shape=(3, 3), matvec=d.calculate_reaction_force_global)
    assert lattice_of_operator(A) is d

    class Opt(Dummy):
        def op(self):
            def matvec(v):
                return self.calculate_reaction_force_global(v)
            return LinearOperator((3, 3), matvec=matvec, dtype=float)
    o = Opt()
    assert lattice_of_operator(o.op()) is o
    assert lattice_of_operator(LinearOperator((3, 3), matvec=lambda v: v)) is None


@pytest.mark.parametrize("opt_type,extra", [("unit_cell", {"hybrid": False}), ("constant", {"hybrid": False}),
                                            ("linear", {"direction": ["x"]})])
def test_parameter_mapping_equals_the_reference_calculate_gradient(ref, opt_type, extra):
    """Real LatticeOpti objects, the reference's own solve_DDM and finite-difference dS (get_schur_complement bound to
    the oracle): fem.cell_sensitivities_to_parameters(q) with q = u_c^T dS u_c must equal LatticeOpti.calculate_gradient."""
    import contextlib, io
    from oracle import lattice_oracle as orc
    from pylatticedso_b200.fem import cell_sensitivities_to_parameters
    refshim, ls, lo = ref
    opt = {"objective_function": "min", "objective_type": "compliance", "max_iterations": 2,
           "optimization_parameters": dict({"type": opt_type}, **extra), "enable_parameter_normalization": False,
           "simulation_type": "DDM", "enable_gradient_computing": True}
    refshim.set_inline_presets({"o": _cfg((3, 1, 1), opt)})
    orig = ls.get_schur_complement
    ls.get_schur_complement = lambda lattice, cell_index=None: orc.cell_schur_from_lattice(
        lattice, 0 if cell_index is None else cell_index, 1013.0, 0.3, 2)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            lat = lo.LatticeOpti("o")
            lat.enable_normalization = False
            n_par = lat.number_parameters
            theta = np.array([0.05] * n_par) if opt_type != "linear" else np.array([0.004, 0.045])
            lat.objective(theta)
            lat.gradient(theta)
            g_ref = np.asarray(lat.calculate_gradient(), dtype=np.float64)
    finally:
        ls.get_schur_complement = orig
    q = []
    for c in lat.cells:
        u = np.asarray(c.get_displacement_at_nodes(c.node_in_order_simulation), dtype=np.float64).ravel()
        q.append([float(u @ (dS @ u)) for dS in c.schur_complement_gradient])
    got = cell_sensitivities_to_parameters(lat, np.array(q))
    assert got.shape == g_ref.shape and np.any(g_ref != 0)
    assert np.abs(got - g_ref).max() <= 1e-12 * np.abs(g_ref).max()
