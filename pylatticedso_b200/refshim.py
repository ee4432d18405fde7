"""Import shim that lets the *unmodified* host-side Python of pyLatticeDSO
(``pyLatticeDesign`` geometry/BC bookkeeping, ``pyLatticeSim.lattice_sim``,
``pyLatticeOpti.lattice_opti``) be imported on a machine that has none of the
FEniCSx / gmsh / PETSc stack installed.

The B200 hot path replaces everything those packages did numerically
(SURVEY.md section 8), so the only thing still needed from them is that the
``import`` statements at the top of the reference modules succeed
(``pyLatticeDesign/lattice.py:17`` imports gmsh, ``lattice_sim.py:13`` imports
colorama, ...).  Missing packages are replaced by inert stub modules; packages
that *are* installed are left alone.

Nothing here is on the numerical path.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

_STUB_NAMES = [
    "gmsh", "colorama",
    "matplotlib", "matplotlib.colors", "matplotlib.pyplot", "matplotlib.cm",
    "matplotlib.widgets", "matplotlib.patches", "matplotlib.lines",
    "matplotlib.collections", "matplotlib.figure", "matplotlib.axes",
    "matplotlib.ticker", "matplotlib.animation",
    "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
    "ufl", "basix", "basix.ufl",
    "dolfinx", "dolfinx.fem", "dolfinx.fem.petsc", "dolfinx.io", "dolfinx.io.gmshio",
    "dolfinx.mesh", "dolfinx.common", "dolfinx.geometry",
    "petsc4py", "petsc4py.PETSc", "mpi4py", "mpi4py.MPI", "dolfinx_mpc",
    "trimesh", "rtree", "pyvista",
]


class _Blank:
    """Object whose every attribute is the empty string (colorama.Fore/Style)."""

    def __getattr__(self, _name):
        return ""


class _Stub(types.ModuleType):
    """Module whose attributes are child stubs and which is callable (returns None)."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []  # behave like a package
        self.__file__ = "<pylatticedso_b200.refshim stub>"

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        full = self.__name__ + "." + item
        mod = sys.modules.get(full)
        if mod is None:
            mod = _Stub(full)
            sys.modules[full] = mod
        object.__setattr__(self, item, mod)
        return mod

    def __call__(self, *a, **k):
        return None

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):  # allows ``class X(stub.Something)``
        return (object,)


def _is_missing(name: str) -> bool:
    root = name.split(".")[0]
    if root in sys.modules and not isinstance(sys.modules[root], _Stub):
        return False
    try:
        return importlib.util.find_spec(root) is None
    except (ImportError, ValueError):
        return True


def install_stubs() -> list[str]:
    """Register stub modules for every missing third-party package. Idempotent."""
    done = []
    for name in _STUB_NAMES:
        if name in sys.modules:
            continue
        if not _is_missing(name):
            continue
        mod = _Stub(name)
        if name == "colorama":
            mod.Fore = _Blank()
            mod.Style = _Blank()
            mod.Back = _Blank()
        sys.modules[name] = mod
        parent, _, child = name.rpartition(".")
        if parent and parent in sys.modules:
            object.__setattr__(sys.modules[parent], child, mod)
        done.append(name)
    return done


def reference_root() -> str | None:
    """Directory of a pyLatticeDSO checkout, or None (e.g. on the GPU box)."""
    for cand in (os.environ.get("PYLATTICEDSO_ROOT"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "src", "pyLatticeDesign")):
            return cand
    return None


def have_reference() -> bool:
    return reference_root() is not None


def import_reference():
    """Make ``pyLatticeDesign`` / ``pyLatticeSim`` / ``pyLatticeOpti`` importable.

    Returns the module ``pyLatticeSim.lattice_sim``.  Raises ImportError when no
    checkout is available.
    """
    root = reference_root()
    if root is None:
        raise ImportError("no pyLatticeDSO checkout found (set PYLATTICEDSO_ROOT)")
    install_stubs()
    for p in (os.path.join(root, "src"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("pyLatticeSim.lattice_sim")


def set_inline_presets(presets: dict) -> None:
    """Route ``open_lattice_parameters(name)`` to in-memory dicts.

    The reference resolves JSON presets relative to its own checkout
    (``pyLatticeDesign/utils.py:111-130``); tests and benchmarks want to supply
    configurations without writing into that tree.  Names not in ``presets``
    fall through to the original loader.
    """
    import_reference()
    import pyLatticeDesign.utils as u
    orig = getattr(u, "_b200_orig_open", None) or u.open_lattice_parameters
    u._b200_orig_open = orig

    def _open(name):
        key = str(name)
        if key in presets:
            import copy
            return copy.deepcopy(presets[key])
        return orig(name)

    for modname in ("pyLatticeDesign.utils", "pyLatticeDesign.lattice",
                    "pyLatticeSim.lattice_sim", "pyLatticeOpti.lattice_opti"):
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, "open_lattice_parameters"):
            mod.open_lattice_parameters = _open
