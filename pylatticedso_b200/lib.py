"""ctypes binding of ``liblattice_b200.so`` (the C ABI declared in
``include/lattice_b200.h``).  torch tensors are used only as device buffers:
every call passes ``tensor.data_ptr()``.

There is NO CPU fallback: :func:`load` raises if the shared library is missing
and :class:`Context` raises if no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "liblattice_b200.so")
SOURCES = ["lattice_core.cu", "lattice_solver.cu", "lattice_schur.cu", "lattice_surrogate.cu"]
HEADERS = [os.path.join(_HERE, "csrc", "common.cuh"), os.path.join(_HERE, "csrc", "matfree.cuh"),
           os.path.join(_HERE, "csrc", "pcg_persist.cuh"), os.path.join(_HERE, "csrc", "coarse.cuh"),
           os.path.join(_ROOT, "include", "lattice_b200.h")]

ASM_GATHER, ASM_ATOMIC, ASM_ROWS = 0, 1, 2
PC_NONE, PC_JACOBI, PC_BLOCK6 = 0, 1, 2

EXPORTS = [
    "lat_version", "lat_ctx_create", "lat_ctx_destroy", "lat_last_error", "lat_ctx_sync", "lat_launch_count",
    "lat_elem_stiffness", "lat_bsr_pattern_build", "lat_bsr_pattern_export", "lat_csr_structure",
    "lat_bsr_to_csr_values", "lat_assemble_bsr", "lat_apply_dirichlet", "lat_set_dirichlet_values", "lat_bsr_spmv", "lat_pcg_bsr",
    "lat_matfree_setup", "lat_matfree_apply", "lat_matfree_rhs", "lat_pcg_matfree", "lat_pcg_matfree_dist",
    "lat_compliance_grad", "lat_schur_batch", "lat_schur_batch_chains", "lat_assemble_bsr_struts", "lat_strut_recover", "lat_ddm_matvec",
    "lat_nccl_unique_id", "lat_comm_create", "lat_comm_destroy", "lat_allreduce_sum", "lat_halo_exchange",
    "lat_pcg_bsr_dist", "lat_p2p_arena_create", "lat_p2p_attach", "lat_p2p_destroy", "lat_assemble_cells_bsr",
    "lat_cell_quadform", "lat_schur_batch_struts",
    "lat_greedy_basis", "lat_upper_solve", "lat_basis_project", "lat_rbf_fit", "lat_rbf_eval", "lat_alpha_lookup",
    "lat_basis_prepare", "lat_basis_expand", "lat_alpha_simplex",
    "lat_coarse_setup", "lat_coarse_galerkin", "lat_coarse_set_inverse", "lat_coarse_apply",
    "lat_assemble_cells_bsr_plan",
]


class LatticeB200Error(RuntimeError):
    pass


class PcgOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("mintol", C.c_double), ("alpha_max", C.c_double),
                ("restart_every", C.c_int64), ("maxiter", C.c_int32), ("precond", C.c_int32),
                ("reference_semantics", C.c_int32), ("check_every", C.c_int32),
                ("profile_iters", C.c_int32), ("reserved", C.c_int32)]


class PcgResult(C.Structure):
    _fields_ = [("iters", C.c_int32), ("info", C.c_int32), ("relres", C.c_double), ("norm_b", C.c_double),
                ("solve_ms", C.c_double), ("launches", C.c_int64), ("spmv_ms", C.c_double),
                ("update_ms", C.c_double), ("profiled", C.c_int32), ("reserved", C.c_int32),
                ("true_relres", C.c_double)]


class Halo(C.Structure):
    _fields_ = [("n_neighbors", C.c_int32), ("pad", C.c_int32), ("peer", C.POINTER(C.c_int32)),
                ("send_count", C.POINTER(C.c_int32)), ("recv_count", C.POINTER(C.c_int32)),
                ("send_idx", C.c_void_p), ("n_owned", C.c_int64), ("n_local", C.c_int64)]


def nvcc_command(out=LIB_PATH):
    src = [os.path.join(_HERE, "csrc", s) for s in SOURCES]
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(_ROOT, "include"), "-o", out] + src + ["-ldl"]


def build(force=False, verbose=False):
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    src = [os.path.join(_HERE, "csrc", s) for s in SOURCES] + HEADERS
    if not force and os.path.exists(LIB_PATH):
        t = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(s) <= t for s in src):
            return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise LatticeB200Error("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def load():
    """dlopen the library and declare the prototypes. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LatticeB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA path is the only path; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.lat_version.restype = C.c_int
    lib.lat_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.lat_ctx_destroy.argtypes = [vp]
    lib.lat_last_error.argtypes = [vp]
    lib.lat_last_error.restype = C.c_char_p
    lib.lat_ctx_sync.argtypes = [vp]
    lib.lat_launch_count.argtypes = [vp]
    lib.lat_launch_count.restype = i64
    lib.lat_elem_stiffness.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, C.c_int, vp]
    lib.lat_bsr_pattern_build.argtypes = [vp, vp, vp, i64, i64, C.POINTER(i64)]
    lib.lat_bsr_pattern_export.argtypes = [vp, vp, vp, vp]
    lib.lat_csr_structure.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.lat_bsr_to_csr_values.argtypes = [vp, vp, i64, vp, vp]
    lib.lat_assemble_bsr.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, dbl, dbl, dbl, C.c_int, C.c_int, vp]
    lib.lat_apply_dirichlet.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
    lib.lat_set_dirichlet_values.argtypes = [vp, vp, vp, i64, vp]
    lib.lat_bsr_spmv.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    lib.lat_pcg_bsr.argtypes = [vp, vp, vp, vp, i64, vp, vp, C.POINTER(PcgOpts), C.POINTER(PcgResult)]
    lib.lat_matfree_setup.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, C.c_double, C.c_double, C.c_double, vp]
    lib.lat_matfree_apply.argtypes = [vp, vp, vp, C.c_int]
    lib.lat_matfree_rhs.argtypes = [vp, vp, vp, vp]
    lib.lat_pcg_matfree.argtypes = [vp, vp, vp, C.POINTER(PcgOpts), C.POINTER(PcgResult)]
    lib.lat_pcg_matfree_dist.argtypes = [vp, C.POINTER(Halo), vp, vp, C.POINTER(PcgOpts), C.POINTER(PcgResult)]
    lib.lat_compliance_grad.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, vp, vp, i64, vp, vp]
    lib.lat_schur_batch.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, i32, dbl, dbl, dbl, vp, vp, vp, i32, vp]
    lib.lat_strut_recover.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, dbl, dbl, dbl, vp, vp]
    lib.lat_assemble_bsr_struts.argtypes = [vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, i64, i64, dbl, dbl, dbl, vp]
    lib.lat_schur_batch_chains.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, dbl, dbl, dbl, vp]
    lib.lat_schur_batch_struts.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, dbl, dbl, dbl, vp,
                                           vp, vp, i32, vp]
    lib.lat_ddm_matvec.argtypes = [vp, vp, i64, vp, vp, i64, i32, i64, vp, vp]
    lib.lat_assemble_cells_bsr.argtypes = [vp, vp, i64, vp, i64, i32, vp, vp, i64, vp]
    lib.lat_assemble_cells_bsr_plan.argtypes = [vp, vp, i64, i32, vp, vp, i64, vp]
    lib.lat_cell_quadform.argtypes = [vp, vp, i64, vp, vp, vp, i64, i32, i32, vp]
    lib.lat_nccl_unique_id.argtypes = [vp]
    lib.lat_comm_create.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.lat_comm_destroy.argtypes = [vp]
    lib.lat_allreduce_sum.argtypes = [vp, vp, i64]
    lib.lat_halo_exchange.argtypes = [vp, C.POINTER(Halo), vp]
    lib.lat_p2p_arena_create.argtypes = [vp, i64, vp]
    lib.lat_p2p_attach.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.lat_p2p_destroy.argtypes = [vp]
    lib.lat_pcg_bsr_dist.argtypes = [vp, vp, vp, vp, C.POINTER(Halo), vp, vp, C.POINTER(PcgOpts), C.POINTER(PcgResult)]
    lib.lat_greedy_basis.argtypes = [vp, vp, i64, i64, dbl, vp, vp, vp, vp, C.POINTER(i32)]
    lib.lat_upper_solve.argtypes = [vp, vp, i32, i32, vp, i32]
    lib.lat_basis_project.argtypes = [vp, vp, i32, i64, vp, i64, vp]
    lib.lat_rbf_fit.argtypes = [vp, vp, i32, i32, vp, i32, dbl, vp]
    lib.lat_rbf_eval.argtypes = [vp, vp, i32, i32, vp, i32, vp, i64, vp, vp]
    lib.lat_alpha_lookup.argtypes = [vp, i32, vp, i32, i32, vp, i32, vp, i64, vp]
    lib.lat_alpha_simplex.argtypes = [vp, vp, vp, i32, i32, vp, i32, vp, i32, vp, i64, vp]
    lib.lat_basis_prepare.argtypes = [vp, vp, i64, i32, i32, vp]
    lib.lat_basis_expand.argtypes = [vp, vp, i32, i64, vp, i64, i32, vp]
    lib.lat_coarse_setup.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, i32, vp, vp]
    lib.lat_coarse_galerkin.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.lat_coarse_set_inverse.argtypes = [vp, vp]
    lib.lat_coarse_apply.argtypes = [vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("lat_last_error", "lat_launch_count"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


_default_ctx = {}


def default_context(device=None):
    """Process-wide context of a device (created on first use): what every host entry point uses when the caller
    passes ``ctx=None``, so that repeated drop-in calls share one workspace instead of allocating a context each."""
    import torch
    if not torch.cuda.is_available():
        raise LatticeB200Error("no CUDA device available: pylatticedso_b200 has no CPU fallback")
    idx = torch.cuda.current_device() if device is None else (device if isinstance(device, int) else torch.device(device).index or 0)
    c = _default_ctx.get(idx)
    if c is None or c.h is None:
        c = Context(idx)
        _default_ctx[idx] = c
    return c


class Context:
    """One library context = one GPU + one CUDA stream (torch's current stream)."""

    def __init__(self, device=None):
        import torch
        self.lib = load()
        if not torch.cuda.is_available():
            raise LatticeB200Error("no CUDA device available: pylatticedso_b200 has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.current_stream(self.device)
        h = C.c_void_p()
        rc = self.lib.lat_ctx_create(self.device.index, C.c_void_p(self.stream.cuda_stream), C.byref(h))
        if rc != 0:
            raise LatticeB200Error(f"lat_ctx_create failed with code {rc}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.lat_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            msg = self.lib.lat_last_error(self.h)
            raise LatticeB200Error(f"[{rc}] {msg.decode() if msg else 'unknown error'}")

    def sync(self):
        self.check(self.lib.lat_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.lat_launch_count(self.h))

    # -- thin typed wrappers (device tensors in, device tensors out) ----------
    def elem_stiffness(self, x, y, z, en0, en1, rad, young, nu, kappa=0.9, drad=False):
        import torch
        ne = en0.numel()
        Ke = torch.empty((ne, 12, 12), dtype=torch.float64, device=self.device)
        self.check(self.lib.lat_elem_stiffness(self.h, _ptr(x), _ptr(y), _ptr(z), _ptr(en0), _ptr(en1), _ptr(rad),
                                               ne, young, nu, kappa, int(drad), _ptr(Ke)))
        return Ke

    def bsr_pattern(self, en0, en1, n_nodes, want_elem_block=False):
        import torch
        nnzb = C.c_int64(0)
        self.check(self.lib.lat_bsr_pattern_build(self.h, _ptr(en0), _ptr(en1), en0.numel(), n_nodes, C.byref(nnzb)))
        rowptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=self.device)
        colidx = torch.empty(nnzb.value, dtype=torch.int32, device=self.device)
        eb = torch.empty((en0.numel(), 4), dtype=torch.int32, device=self.device) if want_elem_block else None
        self.check(self.lib.lat_bsr_pattern_export(self.h, _ptr(rowptr), _ptr(colidx), _ptr(eb)))
        return (rowptr, colidx, eb) if want_elem_block else (rowptr, colidx)

    def csr_structure(self, rowptr, colidx):
        import torch
        n_nodes = rowptr.numel() - 1
        indptr = torch.empty(6 * n_nodes + 1, dtype=torch.int32, device=self.device)
        indices = torch.empty(36 * colidx.numel(), dtype=torch.int32, device=self.device)
        self.check(self.lib.lat_csr_structure(self.h, _ptr(rowptr), _ptr(colidx), n_nodes, _ptr(indptr), _ptr(indices)))
        return indptr, indices

    def bsr_to_csr_values(self, rowptr, vals):
        import torch
        out = torch.empty_like(vals)
        self.check(self.lib.lat_bsr_to_csr_values(self.h, _ptr(rowptr), rowptr.numel() - 1, _ptr(vals), _ptr(out)))
        return out

    def assemble_bsr(self, x, y, z, en0, en1, rad, n_nodes, nnzb, young, nu, kappa=0.9, mode=ASM_ROWS,
                     drad=False, chain=None, out=None):
        import torch
        vals = out if out is not None else torch.empty(nnzb * 36, dtype=torch.float64, device=self.device)
        self.check(self.lib.lat_assemble_bsr(self.h, _ptr(x), _ptr(y), _ptr(z), _ptr(en0), _ptr(en1), _ptr(rad),
                                             _ptr(chain), en0.numel(), n_nodes, young, nu, kappa, mode, int(drad),
                                             _ptr(vals)))
        return vals

    def apply_dirichlet(self, rowptr, colidx, vals, fixed, g, f, inplace=False, want_matrix=True):
        import torch
        n_nodes = rowptr.numel() - 1
        b = torch.empty(6 * n_nodes, dtype=torch.float64, device=self.device)
        vbc = (vals if inplace else torch.empty_like(vals)) if want_matrix else None
        self.check(self.lib.lat_apply_dirichlet(self.h, _ptr(rowptr), _ptr(colidx), n_nodes, _ptr(vals), _ptr(fixed),
                                                _ptr(g), _ptr(f), _ptr(vbc), _ptr(b)))
        return vbc, b

    def set_dirichlet_values(self, fixed, g, u):
        self.check(self.lib.lat_set_dirichlet_values(self.h, _ptr(fixed), _ptr(g), u.numel(), _ptr(u)))
        return u

    def spmv(self, rowptr, colidx, vals, x, out=None):
        import torch
        y = out if out is not None else torch.empty_like(x)
        self.check(self.lib.lat_bsr_spmv(self.h, _ptr(rowptr), _ptr(colidx), _ptr(vals), rowptr.numel() - 1, _ptr(x), _ptr(y)))
        return y

    def pcg(self, rowptr, colidx, vals, b, x=None, tol=1e-8, maxiter=10000, precond=PC_JACOBI,
            reference_semantics=False, mintol=0.0, alpha_max=0.0, restart_every=0, check_every=0,
            profile_iters=0, tma_spmv=False, classic=False, debug=0, persistent=True):
        """``persistent=True`` (default): systems whose vectors fit the GPU's shared memory are solved by ONE
        persistent cooperative kernel (csrc/pcg_persist.cuh; ``info['persistent']``); larger ones, reference
        semantics and profiled runs use the three-kernel iteration."""
        import torch
        if x is None:
            x = torch.empty_like(b)
        o = PcgOpts(tol, mintol, alpha_max, restart_every, maxiter, precond, int(reference_semantics), check_every,
                    profile_iters, (2 if tma_spmv else 0) | (8 if classic else 0) | (0 if persistent else 128) | (debug << 8))
        r = PcgResult()
        self.check(self.lib.lat_pcg_bsr(self.h, _ptr(rowptr), _ptr(colidx), _ptr(vals), rowptr.numel() - 1, _ptr(b),
                                        _ptr(x), C.byref(o), C.byref(r)))
        return x, dict(iters=r.iters, info=r.info, relres=r.relres, norm_b=r.norm_b, solve_ms=r.solve_ms,
                       launches=r.launches, spmv_ms=r.spmv_ms, update_ms=r.update_ms, profiled=r.profiled,
                       true_relres=r.true_relres, restarts=r.reserved & 0xff, graph=bool(r.reserved & 0x100),
                       persistent=bool(r.reserved & 0x200), two_level=bool(r.reserved & 0x400))

    # ---- matrix-free operator (resident in the context; needs bsr_pattern() of the same mesh) ----
    def matfree_setup(self, x, y, z, en0, en1, rad, n_nodes, young, nu, kappa=0.9, fixed=None):
        self.check(self.lib.lat_matfree_setup(self.h, _ptr(x), _ptr(y), _ptr(z), _ptr(en0), _ptr(en1), _ptr(rad),
                                              en0.numel(), n_nodes, young, nu, kappa, _ptr(fixed)))

    def matfree_apply(self, u, out=None, eliminated=True):
        import torch
        if out is None:
            out = torch.empty_like(u)
        self.check(self.lib.lat_matfree_apply(self.h, _ptr(u), _ptr(out), int(eliminated)))
        return out

    def matfree_rhs(self, g, f=None, out=None):
        import torch
        if out is None:
            out = torch.empty_like(g)
        self.check(self.lib.lat_matfree_rhs(self.h, _ptr(g), _ptr(f), _ptr(out)))
        return out

    def pcg_matfree(self, b, x=None, tol=1e-8, maxiter=10000, precond=PC_JACOBI, check_every=0, profile_iters=0):
        import torch
        if x is None:
            x = torch.empty_like(b)
        o = PcgOpts(tol, 0.0, 0.0, 0, maxiter, precond, 0, check_every, profile_iters, 0)
        r = PcgResult()
        self.check(self.lib.lat_pcg_matfree(self.h, _ptr(b), _ptr(x), C.byref(o), C.byref(r)))
        return x, dict(iters=r.iters, info=r.info, relres=r.relres, norm_b=r.norm_b, solve_ms=r.solve_ms,
                       launches=r.launches, spmv_ms=r.spmv_ms, update_ms=r.update_ms, profiled=r.profiled,
                       true_relres=r.true_relres, restarts=r.reserved & 0xff, graph=bool(r.reserved & 0x100),
                       two_level=bool(r.reserved & 0x400))

    # ---- two-level preconditioner (csrc/coarse.cuh; host policy in coarse.py) ----
    def coarse_setup(self, x, y, z, node_agg, agg_ptr, agg_nodes, fixed=None, centers=None):
        self._coarse_keep = None
        self._coarse_owner = None
        self.check(self.lib.lat_coarse_setup(self.h, _ptr(x), _ptr(y), _ptr(z), x.numel(), _ptr(node_agg), _ptr(agg_ptr),
                                             _ptr(agg_nodes), agg_ptr.numel() - 1, _ptr(fixed), _ptr(centers)))

    def coarse_galerkin(self, rowptr, colidx, vals, n_agg, n_rows=None):
        import torch
        E = torch.empty((6 * n_agg, 6 * n_agg), dtype=torch.float64, device=self.device)
        n_rows = rowptr.numel() - 1 if n_rows is None else int(n_rows)
        self.check(self.lib.lat_coarse_galerkin(self.h, _ptr(rowptr), _ptr(colidx), _ptr(vals), n_rows, _ptr(E)))
        return E

    def coarse_set_inverse(self, einv):
        """Registers (and keeps alive) the dense coarse inverse; ``None`` switches the coarse correction off."""
        self._coarse_keep = einv
        self.check(self.lib.lat_coarse_set_inverse(self.h, _ptr(einv)))

    def coarse_apply(self, r, u):
        self.check(self.lib.lat_coarse_apply(self.h, _ptr(r), _ptr(u)))
        return u

    def compliance_grad(self, x, y, z, en0, en1, rad, group, n_groups, u, young, nu, kappa=0.9, chain=None,
                        lam=None, want_elem=False):
        import torch
        g = torch.empty(n_groups, dtype=torch.float64, device=self.device)
        q = torch.empty(en0.numel(), dtype=torch.float64, device=self.device) if want_elem else None
        self.check(self.lib.lat_compliance_grad(self.h, _ptr(x), _ptr(y), _ptr(z), _ptr(en0), _ptr(en1), _ptr(rad),
                                                _ptr(chain), _ptr(group), en0.numel(), young, nu, kappa, _ptr(u),
                                                _ptr(lam), n_groups, _ptr(g), _ptr(q)))
        return (g, q) if want_elem else g

    def schur_batch(self, xyz, len0, len1, rad, n_bnd_nodes, young, nu, kappa=0.9, elem_group=None, chain=None,
                    n_grad=0):
        """xyz [n_cells, n_loc_nodes, 3], rad [n_cells, n_loc_elem] -> S [n_cells, nB, nB] (and dS)."""
        import torch
        n_cells, nn = int(xyz.shape[0]), int(xyz.shape[1])
        ne = int(len0.numel())
        nB = 6 * n_bnd_nodes
        S = torch.empty((n_cells, nB, nB), dtype=torch.float64, device=self.device)
        dS = torch.empty((n_cells, n_grad, nB, nB), dtype=torch.float64, device=self.device) if n_grad > 0 else None
        self.check(self.lib.lat_schur_batch(self.h, _ptr(xyz), _ptr(len0), _ptr(len1), _ptr(rad), n_cells, nn,
                                            n_bnd_nodes, ne, young, nu, kappa, _ptr(S), _ptr(elem_group), _ptr(chain),
                                            n_grad, _ptr(dS)))
        return (S, dS) if n_grad > 0 else S

    def assemble_bsr_struts(self, xyz, len0, len1, rad, chain_ptr, chain_elem, chain_flip, n_joints, nnzb, young, nu,
                            kappa=0.9, out=None):
        """Joint-only BSR values (the pattern of the joint mesh must be resident: bsr_pattern(chain_a, chain_b, n_joints))."""
        import torch
        if out is None:
            out = torch.empty(nnzb * 36, dtype=torch.float64, device=self.device)
        self.check(self.lib.lat_assemble_bsr_struts(self.h, _ptr(xyz), _ptr(len0), _ptr(len1), _ptr(rad), int(xyz.shape[0]),
                                                    int(len0.numel()), _ptr(chain_ptr), _ptr(chain_elem), _ptr(chain_flip),
                                                    int(chain_ptr.numel()) - 1, n_joints, young, nu, kappa, _ptr(out)))
        return out

    def strut_recover(self, xyz, len0, len1, rad, chain_ptr, chain_elem, chain_flip, chain_a, chain_b, max_len, young, nu,
                      kappa, u_joints, u_full):
        self.check(self.lib.lat_strut_recover(self.h, _ptr(xyz), _ptr(len0), _ptr(len1), _ptr(rad), _ptr(chain_ptr),
                                              _ptr(chain_elem), _ptr(chain_flip), _ptr(chain_a), _ptr(chain_b),
                                              int(chain_a.numel()), int(max_len), young, nu, kappa, _ptr(u_joints), _ptr(u_full)))
        return u_full

    def schur_batch_chains(self, xyz, len0, len1, rad, chains, n_bnd_nodes, young, nu, kappa=0.9):
        """Schur complements through the strut pre-pass.  ``chains``: dict of device int32 tensors
        (ptr, elem, flip, a, b) + n_joints, see ``schur.strut_chains``."""
        import torch
        n_cells, nn = int(xyz.shape[0]), int(xyz.shape[1])
        nB = 6 * n_bnd_nodes
        S = torch.empty((n_cells, nB, nB), dtype=torch.float64, device=self.device)
        self.check(self.lib.lat_schur_batch_chains(self.h, _ptr(xyz), _ptr(len0), _ptr(len1), _ptr(rad), n_cells, nn,
                                                   int(len0.numel()), _ptr(chains["ptr"]), _ptr(chains["elem"]),
                                                   _ptr(chains["flip"]), _ptr(chains["a"]), _ptr(chains["b"]),
                                                   int(chains["a"].numel()), int(chains["n_joints"]), n_bnd_nodes,
                                                   young, nu, kappa, _ptr(S)))
        return S

    def schur_batch_struts(self, xyz, len0, len1, rad, chains, n_bnd_nodes, young, nu, kappa=0.9, chain_group=None,
                           drad_chain=None, n_grad=0):
        """Schur complements (and dS for ``n_grad`` radius groups) through the strut pre-pass; star cells (BCC) run in
        the half-warp kernel, other topologies fall through to ``schur_batch_chains`` (values only)."""
        import torch
        n_cells, nn = int(xyz.shape[0]), int(xyz.shape[1])
        nB = 6 * n_bnd_nodes
        S = torch.empty((n_cells, nB, nB), dtype=torch.float64, device=self.device)
        dS = torch.empty((n_cells, n_grad, nB, nB), dtype=torch.float64, device=self.device) if n_grad > 0 else None
        self.check(self.lib.lat_schur_batch_struts(self.h, _ptr(xyz), _ptr(len0), _ptr(len1), _ptr(rad), n_cells, nn,
                                                   int(len0.numel()), _ptr(chains["ptr"]), _ptr(chains["elem"]),
                                                   _ptr(chains["flip"]), _ptr(chains["a"]), _ptr(chains["b"]),
                                                   int(chains["a"].numel()), int(chains["n_joints"]), n_bnd_nodes,
                                                   young, nu, kappa, _ptr(S), _ptr(chain_group), _ptr(drad_chain),
                                                   n_grad, _ptr(dS)))
        return (S, dS) if n_grad > 0 else S

    def ddm_matvec(self, S, gidx, x, n_free=None, u_fixed=None, out=None):
        """y = sum_c B_c S_c B_c^T x.  S: [n_cells, nb, nb] or [nb, nb] (shared by all cells)."""
        import torch
        n_cells, nb = int(gidx.shape[0]), int(gidx.shape[1])
        stride = 0 if S.dim() == 2 else nb * nb
        n_free = int(x.numel()) if n_free is None else n_free
        y = out if out is not None else torch.empty(n_free, dtype=torch.float64, device=self.device)
        self.check(self.lib.lat_ddm_matvec(self.h, _ptr(S), stride, _ptr(gidx), _ptr(u_fixed), n_cells, nb, n_free,
                                           _ptr(x), _ptr(y)))
        return y

    def assemble_cells_bsr(self, S, cell_nodes, rowptr, colidx):
        """K_G = sum_c P_c^T S_c P_c.  S: [n_cells, nb, nb] or [nb, nb] (shared)."""
        import torch
        n_cells, nbn = int(cell_nodes.shape[0]), int(cell_nodes.shape[1])
        stride = 0 if S.dim() == 2 else 36 * nbn * nbn
        vals = torch.empty(colidx.numel() * 36, dtype=torch.float64, device=self.device)
        self.check(self.lib.lat_assemble_cells_bsr(self.h, _ptr(S), stride, _ptr(cell_nodes), n_cells, nbn, _ptr(rowptr),
                                                   _ptr(colidx), colidx.numel(), _ptr(vals)))
        return vals

    def assemble_cells_plan(self, S, n_bnd_nodes, blk_ptr, contrib, out=None):
        """The same matrix by the plan-driven gather (``lat_assemble_cells_bsr_plan``): fixed summation order."""
        import torch
        nnzb = int(blk_ptr.numel()) - 1
        stride = 0 if S.dim() == 2 else 36 * n_bnd_nodes * n_bnd_nodes
        vals = out if out is not None else torch.empty(nnzb * 36, dtype=torch.float64, device=self.device)
        self.check(self.lib.lat_assemble_cells_bsr_plan(self.h, _ptr(S), stride, n_bnd_nodes, _ptr(blk_ptr), _ptr(contrib),
                                                        nnzb, _ptr(vals)))
        return vals

    def cell_quadform(self, mats, mat_index, U, V=None):
        """q[c, j] = V[c]^T mats[mat_index[c, j]] U[c]  (V = None: V = U)."""
        import torch
        n_cells, n_grad = int(mat_index.shape[0]), int(mat_index.shape[1])
        nb = int(U.shape[1])
        out = torch.empty((n_cells, n_grad), dtype=torch.float64, device=self.device)
        self.check(self.lib.lat_cell_quadform(self.h, _ptr(mats), int(mats.shape[0]), _ptr(mat_index), _ptr(U), _ptr(V),
                                              n_cells, n_grad, nb, _ptr(out)))
        return out

    # -- multi-GPU ---------------------------------------------------------------
    def comm_create(self, rank, world):
        """Create the library's NCCL communicator; the 128-byte id travels through torch.distributed."""
        import torch
        import torch.distributed as dist
        buf = (C.c_char * 128)()
        if rank == 0:
            rc = self.lib.lat_nccl_unique_id(buf)
            if rc != 0:
                raise LatticeB200Error(f"lat_nccl_unique_id failed with code {rc}")
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.to(self.device)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().numpy().tobytes())
        idbuf = C.create_string_buffer(raw, 128)
        self.check(self.lib.lat_comm_create(self.h, idbuf, world, rank))
        self.rank, self.world = rank, world

    def comm_destroy(self):
        self.lib.lat_comm_destroy(self.h)

    def make_halo(self, peers, send_counts, recv_counts, send_idx_dev, n_owned, n_local):
        n = len(peers)
        arr = lambda v: (C.c_int32 * max(1, n))(*[int(q) for q in v])
        keep = (arr(peers), arr(send_counts), arr(recv_counts), send_idx_dev)
        h = Halo(n, 0, C.cast(keep[0], C.POINTER(C.c_int32)), C.cast(keep[1], C.POINTER(C.c_int32)),
                 C.cast(keep[2], C.POINTER(C.c_int32)), C.c_void_p(send_idx_dev.data_ptr() if send_idx_dev is not None and send_idx_dev.numel() else 0),
                 int(n_owned), int(n_local))
        h._keep = keep
        return h

    def halo_exchange(self, halo, vec):
        self.check(self.lib.lat_halo_exchange(self.h, C.byref(halo), _ptr(vec)))
        return vec

    def allreduce_sum(self, t):
        self.check(self.lib.lat_allreduce_sum(self.h, _ptr(t), t.numel()))
        return t

    def p2p_setup(self, n_local, peers, dst_node0):
        """Create this rank's peer-memory arena, all-gather the IPC handles and map every rank's arena."""
        import torch
        import torch.distributed as dist
        buf = (C.c_char * 64)()
        self.check(self.lib.lat_p2p_arena_create(self.h, int(n_local), buf))
        mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            mine = mine.to(self.device)
        allh = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine)
        raw = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        n = len(peers)
        nb = (C.c_int32 * max(1, n))(*[int(q) for q in peers])
        d0 = (C.c_int64 * max(1, n))(*[int(q) for q in dst_node0])
        self.check(self.lib.lat_p2p_attach(self.h, C.create_string_buffer(raw, len(raw)), self.world, self.rank, n, nb, d0))
        self.p2p_ready = True

    def p2p_destroy(self):
        self.lib.lat_p2p_destroy(self.h)
        self.p2p_ready = False

    def pcg_matfree_dist(self, halo, b, x, tol=1e-8, maxiter=10000, precond=PC_JACOBI, check_every=0, p2p=False,
                         no_graph=False, profile_iters=0, overlap=False, fused_halo=True):
        o = PcgOpts(tol, 0.0, 0.0, 0, maxiter, precond, 0, check_every, profile_iters,
                    (16 if p2p else 0) | (4 if no_graph else 0) | (32 if overlap else 0) | (0 if fused_halo else 64))
        r = PcgResult()
        self.check(self.lib.lat_pcg_matfree_dist(self.h, C.byref(halo), _ptr(b), _ptr(x), C.byref(o), C.byref(r)))
        return x, dict(iters=r.iters, info=r.info, relres=r.relres, norm_b=r.norm_b, solve_ms=r.solve_ms,
                       launches=r.launches, true_relres=r.true_relres, restarts=r.reserved & 0xff, graph=bool(r.reserved & 0x100), spmv_ms=r.spmv_ms,
                       update_ms=r.update_ms, profiled=r.profiled, two_level=bool(r.reserved & 0x400))

    def pcg_dist(self, rowptr, colidx, vals, halo, b, x, tol=1e-8, maxiter=10000, precond=PC_JACOBI,
                 reference_semantics=False, mintol=0.0, alpha_max=0.0, restart_every=0, check_every=0, p2p=False,
                 no_graph=False, profile_iters=0, overlap=False, fused_halo=True, persistent=True):
        """``persistent=True`` (default, peer-memory path with the fused halo): slabs whose vectors fit every GPU's
        shared memory are solved by one persistent cooperative kernel per rank with the halo exchange and the
        all-reduce inside its grid barriers (csrc/pcg_persist.cuh, "Multi-GPU"; ``info['persistent']``)."""
        o = PcgOpts(tol, mintol, alpha_max, restart_every, maxiter, precond, int(reference_semantics), check_every,
                    profile_iters,
                    (16 if p2p else 0) | (4 if no_graph else 0) | (32 if overlap else 0) | (0 if fused_halo else 64) |
                    (0 if persistent else 128))
        r = PcgResult()
        self.check(self.lib.lat_pcg_bsr_dist(self.h, _ptr(rowptr), _ptr(colidx), _ptr(vals), C.byref(halo), _ptr(b),
                                             _ptr(x), C.byref(o), C.byref(r)))
        return x, dict(iters=r.iters, info=r.info, relres=r.relres, norm_b=r.norm_b, solve_ms=r.solve_ms,
                       launches=r.launches, true_relres=r.true_relres, restarts=r.reserved & 0xff, graph=bool(r.reserved & 0x100), spmv_ms=r.spmv_ms,
                       update_ms=r.update_ms, profiled=r.profiled, persistent=bool(r.reserved & 0x200),
                       two_level=bool(r.reserved & 0x400))
