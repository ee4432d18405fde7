"""Rebind the reference's three module-level seams to the B200 path.

    from pylatticedso_b200 import install
    install.patch_reference()          # after `import pyLatticeSim...`

The reference resolves ``solve_FEM_FenicsX``, ``get_schur_complement`` and
``conjugate_gradient_solver`` as module-level names (``lattice_opti.py:22-23``,
``lattice_sim.py:21-22``), so no reference file has to change.
"""
from __future__ import annotations

import sys


def patch_reference(elements_per_strut="gmsh"):
    from . import fem, schur
    done = []

    def _solve(lattice):
        return fem.solve_FEM_B200(lattice, elements_per_strut=elements_per_strut)

    def _schur(lattice, cell_index=None):
        return schur.get_schur_complement(lattice, cell_index, elements_per_strut=elements_per_strut)

    # conjugate_gradient_solver is NOT rebound blindly: the reference passes scipy LinearOperators wrapping its
    # Python cell loop (lattice_sim.py:1148-1160); the device version (pylatticedso_b200.pcg) takes a BsrOperator.
    # The DDM path is replaced as a whole by ddm.solve_DDM_B200 instead.
    for modname, attr, fn in (("pyLatticeSim.utils_simulation", "solve_FEM_FenicsX", _solve),
                              ("pyLatticeOpti.lattice_opti", "solve_FEM_FenicsX", _solve),
                              ("pyLatticeSim.utils_schur", "get_schur_complement", _schur),
                              ("pyLatticeSim.lattice_sim", "get_schur_complement", _schur)):
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, attr):
            setattr(mod, attr, fn)
            done.append(f"{modname}.{attr}")
    return done
