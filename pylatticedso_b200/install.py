"""Rebind the reference's seams to the B200 path.

    from pylatticedso_b200 import install
    install.patch_reference()          # after `import pyLatticeSim...` / `import pyLatticeOpti...`

The reference has no plugin layer; its hot path is reached through module-level names and methods, all of which can
be rebound without touching a reference file:

| name (module / class)                                                    | reference            | replaced by |
|---|---|---|
| ``solve_FEM_FenicsX`` (utils_simulation, lattice_opti)                     | utils_simulation.py:21-56 | ``fem.solve_FEM_B200`` |
| ``get_schur_complement`` (utils_schur, lattice_sim)                        | utils_schur.py:22-53      | ``schur.get_schur_complement`` |
| ``conjugate_gradient_solver`` (conjugate_gradient_solver, lattice_sim, lattice_opti) | conjugate_gradient_solver.py:15-122 | ``pcg.conjugate_gradient_solver`` (recognises the LinearOperator built around ``calculate_reaction_force_global``) |
| ``LatticeSim.solve_DDM``                                                   | lattice_sim.py:1111-1176  | ``ddm.solve_DDM_B200`` |
| ``LatticeSim._compute_schur_gradients``                                    | lattice_sim.py:1020-1054  | ``schur.schur_gradients`` (analytic, no finite differences) |
| ``LatticeOpti.calculate_gradient`` (compliance branch)                     | lattice_opti.py:735-841   | ``ddm.compliance_gradient_cells`` + ``fem.cell_sensitivities_to_parameters`` |
| ``reduce_basis_greedy``, ``project_to_reduced_basis`` (greedy_algorithm)  | greedy_algorithm.py:35-155, 233-266 | ``surrogate.reduce_basis_greedy`` / ``project_to_reduced_basis`` |
| ``ThinPlateSplineRBF`` (utils_rbf, lattice_sim)                            | utils_rbf.py:13-144       | ``surrogate.ThinPlateSplineRBF`` |
| ``LatticeSim.get_schur_complement_from_reduced_basis_batch`` / ``..._from_reduced_basis`` / ``_compute_schur_gradients_RBF`` / ``_define_radial_basis_functions`` | lattice_sim.py:921-1018, 1056-1082, 809-812 | ``surrogate.lattice_*`` (RBF / nearest-neighbour / linear alphas + DMMA ``basis @ alphas``) |

``patch_reference`` returns the list of names it rebound; ``unpatch_reference`` restores the originals.
"""
from __future__ import annotations

import sys

_ORIGINALS = []


def _rebind(owner, attr, fn, done, label):
    if owner is not None and hasattr(owner, attr):
        _ORIGINALS.append((owner, attr, getattr(owner, attr)))
        setattr(owner, attr, fn)
        done.append(label)


def patch_reference(elements_per_strut="gmsh", ctx=None):
    from . import ddm, fem, pcg, schur, surrogate
    done = []

    def _solve(lattice):
        return fem.solve_FEM_B200(lattice, elements_per_strut=elements_per_strut, ctx=ctx)

    def _schur(lattice, cell_index=None):
        return schur.get_schur_complement(lattice, cell_index, elements_per_strut=elements_per_strut, ctx=ctx)

    def _solve_ddm(self):
        return ddm.solve_DDM_B200(self, ctx=ctx)

    def _schur_gradients(self, cell, radii_params):
        return schur.schur_gradients(self, cell, list(radii_params), elements_per_strut=elements_per_strut, ctx=ctx)

    mods = sys.modules
    for modname, attr, fn in (("pyLatticeSim.utils_simulation", "solve_FEM_FenicsX", _solve),
                              ("pyLatticeOpti.lattice_opti", "solve_FEM_FenicsX", _solve),
                              ("pyLatticeSim.utils_schur", "get_schur_complement", _schur),
                              ("pyLatticeSim.lattice_sim", "get_schur_complement", _schur),
                              ("pyLatticeSim.conjugate_gradient_solver", "conjugate_gradient_solver", pcg.conjugate_gradient_solver),
                              ("pyLatticeSim.lattice_sim", "conjugate_gradient_solver", pcg.conjugate_gradient_solver),
                              ("pyLatticeOpti.lattice_opti", "conjugate_gradient_solver", pcg.conjugate_gradient_solver)):
        _rebind(mods.get(modname), attr, fn, done, f"{modname}.{attr}")

    ls = mods.get("pyLatticeSim.lattice_sim")
    lattice_sim_cls = getattr(ls, "LatticeSim", None)
    _rebind(lattice_sim_cls, "solve_DDM", _solve_ddm, done, "pyLatticeSim.lattice_sim.LatticeSim.solve_DDM")
    _rebind(lattice_sim_cls, "_compute_schur_gradients", _schur_gradients, done,
            "pyLatticeSim.lattice_sim.LatticeSim._compute_schur_gradients")

    # N4: reduced-basis / surrogate pipeline
    def _greedy(schur_dict, tol_greedy, file_name=None, verbose=1):
        return surrogate.reduce_basis_greedy(schur_dict, tol_greedy, file_name=file_name, verbose=verbose, ctx=ctx)

    def _project(schur_dict, basis):
        return surrogate.project_to_reduced_basis(schur_dict, basis, ctx=ctx)

    class _TPS(surrogate.ThinPlateSplineRBF):
        def __init__(self, x_train, y_train, reg=0.0):
            super().__init__(x_train, y_train, reg=reg, ctx=ctx)

    for modname in ("pyLatticeSim.greedy_algorithm",):
        _rebind(mods.get(modname), "reduce_basis_greedy", _greedy, done, f"{modname}.reduce_basis_greedy")
        _rebind(mods.get(modname), "project_to_reduced_basis", _project, done, f"{modname}.project_to_reduced_basis")
    for modname in ("pyLatticeSim.utils_rbf", "pyLatticeSim.lattice_sim"):
        _rebind(mods.get(modname), "ThinPlateSplineRBF", _TPS, done, f"{modname}.ThinPlateSplineRBF")
    if lattice_sim_cls is not None and hasattr(lattice_sim_cls, "get_schur_complement_from_reduced_basis_batch"):
        def _batch(self, geometric_params_list):
            return surrogate.lattice_schur_batch(self, geometric_params_list, ctx=ctx)

        def _single(self, geometric_params):
            return surrogate.lattice_schur_single(self, geometric_params, ctx=ctx)

        pre = "pyLatticeSim.lattice_sim.LatticeSim."
        _rebind(lattice_sim_cls, "get_schur_complement_from_reduced_basis_batch", _batch, done,
                pre + "get_schur_complement_from_reduced_basis_batch")
        _rebind(lattice_sim_cls, "get_schur_complement_from_reduced_basis", _single, done,
                pre + "get_schur_complement_from_reduced_basis")
        _rebind(lattice_sim_cls, "_compute_schur_gradients_RBF",
                lambda self, radii_params: surrogate.lattice_schur_gradients_rbf(self, radii_params, ctx=ctx), done,
                pre + "_compute_schur_gradients_RBF")
        _rebind(lattice_sim_cls, "_define_radial_basis_functions", lambda self: surrogate.lattice_define_rbf(self, ctx=ctx), done,
                pre + "_define_radial_basis_functions")

    lo = mods.get("pyLatticeOpti.lattice_opti")
    lattice_opti_cls = getattr(lo, "LatticeOpti", None)
    if lattice_opti_cls is not None and hasattr(lattice_opti_cls, "calculate_gradient"):
        original = lattice_opti_cls.calculate_gradient

        def _calculate_gradient(self):
            if getattr(self, "objective_type", None) != "compliance":
                return original(self)        # adjoint branch: its CG already runs on the device (rebinding above)
            q = ddm.compliance_gradient_cells(self, ctx=ctx)
            return fem.cell_sensitivities_to_parameters(self, q)

        _rebind(lattice_opti_cls, "calculate_gradient", _calculate_gradient, done,
                "pyLatticeOpti.lattice_opti.LatticeOpti.calculate_gradient")
    return done


def unpatch_reference():
    """Restore every name ``patch_reference`` rebound (last in, first out)."""
    n = len(_ORIGINALS)
    while _ORIGINALS:
        owner, attr, fn = _ORIGINALS.pop()
        setattr(owner, attr, fn)
    return n
