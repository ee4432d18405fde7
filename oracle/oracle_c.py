"""ctypes front end of oracle/oracle_c.c (CPU ORACLE, test / baseline infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "oracle_c.c")
LIB = os.path.join(_HERE, "_build", "liboracle_c.so")
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        r = subprocess.run(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed:\n" + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_pcg.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def elem_stiffness(xyz, en, rad, E, nu, kappa=0.9):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); en = np.ascontiguousarray(en, dtype=np.int32)
    rad = np.ascontiguousarray(rad, dtype=np.float64)
    Ke = np.empty((en.shape[0], 12, 12))
    load().orc_elem_stiffness(_p(xyz), _p(en), _p(rad), C.c_long(en.shape[0]), C.c_double(E), C.c_double(nu), C.c_double(kappa), _p(Ke))
    return Ke


def assemble_csr_values(en, Ke, indptr, indices):
    en = np.ascontiguousarray(en, dtype=np.int32); indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.zeros(indices.shape[0])
    load().orc_assemble_csr(C.c_long(en.shape[0]), _p(en), _p(np.ascontiguousarray(Ke)), _p(indptr), _p(indices), _p(data))
    return data


def pcg(indptr, indices, data, b, dinv=None, maxiter=100, tol=1e-5, mintol=1e-5, restart_every=1000, alpha_max=0.1):
    n = b.shape[0]
    x = np.empty(n); it = C.c_int(0)
    indptr = np.ascontiguousarray(indptr, dtype=np.int32); indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    dv = None if dinv is None else np.ascontiguousarray(dinv, dtype=np.float64)
    info = load().orc_pcg(C.c_long(n), _p(indptr), _p(indices), _p(data), _p(b), None if dv is None else _p(dv), _p(x),
                          C.c_int(int(maxiter)), C.c_double(tol), C.c_double(mintol), C.c_long(int(restart_every)),
                          C.c_double(alpha_max), C.byref(it))
    return x, int(info), int(it.value)


def num_threads():
    return int(load().orc_num_threads())


def set_num_threads(n):
    """Explicit OpenMP thread count (overrides an inherited OMP_NUM_THREADS, e.g. torchrun's 1)."""
    load().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def host_threads():
    """Threads this process may run on (cpuset-aware)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1
