"""CPU: the C-ABI library builds for sm_100a, loads, and exports every declared symbol."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            src = open(os.path.join(inc, fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(lat_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol(built_lib):
    from pylatticedso_b200 import lib
    decl = declared_symbols()
    assert decl == set(lib.EXPORTS), decl ^ set(lib.EXPORTS)
    for name in decl:
        assert hasattr(built_lib, name), name
    assert built_lib.lat_version() >= 100


def test_library_contains_sm100a_code(built_lib):
    from pylatticedso_b200 import lib
    out = subprocess.run(["cuobjdump", "--list-elf", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_device_is_a_loud_error(built_lib):
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    assert built_lib.lat_ctx_create(0, None, ctypes.byref(h)) != 0
    from pylatticedso_b200 import lib
    import pytest
    with pytest.raises(lib.LatticeB200Error):
        lib.Context()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pylatticedso_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn
