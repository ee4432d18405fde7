"""Reduced-basis / surrogate Schur pipeline on the GPU.

Host-side mirror of the reference's surrogate-training and surrogate-evaluation code, same names, arguments and
return values, every number computed by ``csrc/lattice_surrogate.cu``:

| reference                                                              | here |
|---|---|
| ``greedy_algorithm.reduce_basis_greedy`` (greedy_algorithm.py:35-155)    | :func:`reduce_basis_greedy` |
| ``greedy_algorithm.project_to_reduced_basis`` (:233-266)                 | :func:`project_to_reduced_basis` |
| ``save_reduced_basis`` / ``load_reduced_basis`` / ``find_name_file_reduced_basis`` (:157-231) | same names (npz schema kept) |
| ``utils_rbf.ThinPlateSplineRBF`` (utils_rbf.py:13-144)                   | :class:`ThinPlateSplineRBF` |
| ``LatticeSim.get_schur_complement_from_reduced_basis_batch`` (lattice_sim.py:921-978), ``..._from_reduced_basis`` (:980-1018), ``_compute_schur_gradients_RBF`` (:1056-1082), ``evaluate_alphas_linear_surrogate`` (:755-807) | :class:`SchurSurrogate` and the ``lattice_*`` drop-ins bound by ``install.patch_reference`` |

There is no CPU fallback: without the CUDA library / a GPU every entry point raises ``LatticeB200Error``.
"""
from __future__ import annotations

import ctypes as C
import re
from math import isqrt
from pathlib import Path

import numpy as np

from . import lib as L

_ptr = L._ptr


def _ctx(ctx):
    return ctx if ctx is not None else L.default_context()


def _dev(ctx, a, dtype=None):
    import torch
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=dtype or np.float64))
    return t.to(ctx.device)


# ---------------------------------------------------------------------------------------------------------------
# greedy reduced basis
# ---------------------------------------------------------------------------------------------------------------
def reduce_basis_greedy(schur_complement_dict_to_reduce: dict, tol_greedy: float, file_name: str = None, verbose: int = 1,
                        ctx=None, save_dir=None):
    """Drop-in for ``reduce_basis_greedy`` (greedy_algorithm.py:35-155): same 7-tuple
    ``(mainelem, reducedcoef, projfieldpp, basis_reduced_ortho, alpha_ortho, matP_sorted, norm_mainelem_sorted)``."""
    import torch
    if not isinstance(schur_complement_dict_to_reduce, dict):
        raise ValueError("schur_complement_dict_to_reduce should be a dict of Schur complements.")      # :85-86
    ctx = _ctx(ctx)
    keys_list = sorted(schur_complement_dict_to_reduce.keys())
    list_elements = np.array(keys_list)
    matrix_schur = np.array([schur_complement_dict_to_reduce[k] for k in keys_list], dtype=np.float64)
    n_snap = matrix_schur.shape[0]
    # vec_F(S) of every snapshot (:99): a transposition of each matrix, done once on the host while staging
    snaps = _dev(ctx, matrix_schur.transpose(0, 2, 1).reshape(n_snap, -1))
    length = snaps.shape[1]
    f64 = dict(dtype=torch.float64, device=ctx.device)
    basis = torch.empty((n_snap, length), **f64)
    coef = torch.zeros((n_snap, n_snap), **f64)
    mainelem_d = torch.zeros(n_snap, dtype=torch.int32, device=ctx.device)
    norms = torch.empty(n_snap, **f64)
    k_out = C.c_int32(0)
    ctx.check(ctx.lib.lat_greedy_basis(ctx.h, _ptr(snaps), n_snap, length, float(tol_greedy), _ptr(basis), _ptr(coef),
                                       _ptr(mainelem_d), _ptr(norms), C.byref(k_out)))
    k = int(k_out.value)
    basis = basis[:k].contiguous()
    mainelem = mainelem_d[:k].cpu().numpy().astype(np.int64)
    # matP = triu(reducedcoef[:, mainelem]); reducedcoef <- matP^-1 reducedcoef (dtrtrs on the device), scaled (:127-130)
    norms_h = norms.cpu().numpy()
    matP_h = np.triu(coef[:k].cpu().numpy()[:, mainelem])
    red = coef[:k].contiguous()
    ctx.check(ctx.lib.lat_upper_solve(ctx.h, _ptr(_dev(ctx, matP_h)), k, k, _ptr(red), n_snap))
    reducedcoef = red.cpu().numpy() * np.outer(1.0 / norms_h[mainelem], norms_h)
    # alpha_ortho[:, S] = lstsq(basis, vec_F(S_S)) (:135-138)
    alphas = torch.empty((n_snap, k), **f64)
    ctx.check(ctx.lib.lat_basis_project(ctx.h, _ptr(basis), k, length, _ptr(snaps), n_snap, _ptr(alphas)))
    vsort = np.argsort(mainelem)
    basis_reduced_ortho = basis.t().contiguous().cpu().numpy()                     # (len, k) as in the reference
    alpha_ortho = alphas.t().contiguous().cpu().numpy()                            # (k, n_snap)
    if file_name is not None:
        save_reduced_basis(file_name, basis_reduced_ortho, alpha_ortho, list_elements, save_dir=save_dir)
    projfieldpp = [matrix_schur[i] for i in mainelem[vsort]]
    if verbose >= 1:
        print("Number of elements in the reduced basis:", len(mainelem))
        print("Selected elements:", list_elements[mainelem[vsort]])
    return (mainelem[vsort], reducedcoef[vsort, :], projfieldpp, basis_reduced_ortho, alpha_ortho,
            matP_h[np.ix_(vsort, vsort)], norms_h[mainelem[vsort]])


def project_to_reduced_basis(schur_complement_dict_to_project: dict, basis_matrix_ortho: np.ndarray, ctx=None):
    """Drop-in for ``project_to_reduced_basis`` (greedy_algorithm.py:233-266): dict key -> alpha vector."""
    import torch
    if not isinstance(schur_complement_dict_to_project, dict):
        raise ValueError("schur_input should be a dict of Schur complements.")
    basis_matrix_ortho = np.asarray(basis_matrix_ortho, dtype=np.float64)
    if basis_matrix_ortho.size == 0:
        raise ValueError("Empty basis_reduced_ortho: build the reduced basis with return_projection_data=True.")
    ctx = _ctx(ctx)
    keys = list(schur_complement_dict_to_project.keys())
    V = _dev(ctx, np.stack([np.ravel(schur_complement_dict_to_project[k_], order="C") for k_ in keys]))     # C order, :253
    B = _dev(ctx, basis_matrix_ortho.T)
    k, length = B.shape
    out = torch.empty((len(keys), k), dtype=torch.float64, device=ctx.device)
    ctx.check(ctx.lib.lat_basis_project(ctx.h, _ptr(B), k, length, _ptr(V), len(keys), _ptr(out)))
    out = out.cpu().numpy()
    return {k_: out[i] for i, k_ in enumerate(keys)}


def find_name_file_reduced_basis(lattice_object_sim, tol_greedy: float):
    """greedy_algorithm.py:211-231."""
    suffix = "_".join(re.sub(r"\W+", "-", str(g)) for g in lattice_object_sim.geom_types)
    tol_str = re.sub(r"e([+-])0+(\d+)$", r"e\1\2", f"{tol_greedy:.0e}")
    return f"reduced_basis_{suffix}_tol_{tol_str}"


def save_reduced_basis(file_name, basis_reduced_ortho, alpha_ortho, list_elements, save_dir=None):
    """npz schema of greedy_algorithm.py:157-184 (``basis_reduced_ortho``, ``alpha_ortho``, ``list_elements``).
    ``save_dir`` replaces the reference's hard-wired ``data/outputs/schur_complement/reduced_basis``."""
    path = Path(save_dir if save_dir is not None else ".") / file_name
    if path.suffix != ".npz":
        path = path.with_suffix(".npz")
    np.savez_compressed(path, basis_reduced_ortho=basis_reduced_ortho, alpha_ortho=alpha_ortho, list_elements=list_elements)
    return path


def load_reduced_basis(path_or_lattice, tol_greedy: float = None, load_dir=None):
    """greedy_algorithm.py:186-209; accepts a path, or (lattice, tol) + the directory that holds the bases."""
    if hasattr(path_or_lattice, "geom_types"):
        path = Path(load_dir if load_dir is not None else ".") / find_name_file_reduced_basis(path_or_lattice, tol_greedy)
    else:
        path = Path(path_or_lattice)
    if path.suffix != ".npz":
        path = path.with_suffix(".npz")
    if not path.is_file():
        raise FileNotFoundError(f"Reduced basis file not found: {path}")
    return np.load(path)


# ---------------------------------------------------------------------------------------------------------------
# thin-plate-spline RBF
# ---------------------------------------------------------------------------------------------------------------
class ThinPlateSplineRBF:
    """Device twin of ``utils_rbf.ThinPlateSplineRBF``: same constructor, ``evaluate`` and ``gradient`` (numpy in /
    numpy out); ``evaluate_device`` / ``gradient_device`` keep the result on the GPU for the batched path."""

    def __init__(self, x_train, y_train, reg: float = 0.0, ctx=None):
        import torch
        self.ctx = _ctx(ctx)
        X = np.asarray(x_train, dtype=float)
        Y = np.asarray(y_train, dtype=float)
        if Y.ndim == 1:
            Y = Y[:, None]                                                       # utils_rbf.py:36-37
        self.x_train, self.y_train = X, Y
        self.N, self.d = X.shape
        _, self.m = Y.shape
        self._X = _dev(self.ctx, X)
        self._wcp = torch.empty((self.N + self.d + 1, self.m), dtype=torch.float64, device=self.ctx.device)
        self.ctx.check(self.ctx.lib.lat_rbf_fit(self.ctx.h, _ptr(self._X), self.N, self.d, _ptr(_dev(self.ctx, Y)), self.m,
                                                float(reg), _ptr(self._wcp)))
        wcp = self._wcp.cpu().numpy()
        self.W, self.CP = wcp[:self.N], wcp[self.N:]

    def _queries(self, x):
        Xq = np.asarray(x, dtype=float)
        single = Xq.ndim == 1
        Xq = Xq[None, :] if single else Xq
        if Xq.shape[1] != self.d:
            raise ValueError(f"queries have {Xq.shape[1]} parameters, the interpolant {self.d}")
        return _dev(self.ctx, Xq), single

    def evaluate_device(self, xq):
        import torch
        M = xq.shape[0]
        f = torch.empty((M, self.m), dtype=torch.float64, device=self.ctx.device)
        self.ctx.check(self.ctx.lib.lat_rbf_eval(self.ctx.h, _ptr(self._X), self.N, self.d, _ptr(self._wcp), self.m, _ptr(xq), M,
                                                 _ptr(f), None))
        return f

    def gradient_device(self, xq):
        import torch
        M = xq.shape[0]
        g = torch.empty((M, self.d, self.m), dtype=torch.float64, device=self.ctx.device)
        self.ctx.check(self.ctx.lib.lat_rbf_eval(self.ctx.h, _ptr(self._X), self.N, self.d, _ptr(self._wcp), self.m, _ptr(xq), M,
                                                 None, _ptr(g)))
        return g

    def evaluate(self, x):
        xq, single = self._queries(x)
        F = self.evaluate_device(xq).cpu().numpy()
        return F[0] if F.shape[0] == 1 else F                                     # utils_rbf.py:104

    def gradient(self, x):
        xq, single = self._queries(x)
        G = self.gradient_device(xq).cpu().numpy()
        return G[0] if G.shape[0] == 1 else G                                     # utils_rbf.py:141


# ---------------------------------------------------------------------------------------------------------------
# surrogate Schur complements
# ---------------------------------------------------------------------------------------------------------------
class SchurSurrogate:
    """Schur complements of many cells from a reduced basis: alpha(parameters) by RBF / nearest neighbour / 1-D linear
    interpolation, then ONE GEMM ``basis @ alphas`` on the FP64 tensor cores (lat_basis_expand), each result in the
    reference's (n, n) layout (order='F' reshape, lattice_sim.py:973-976)."""

    KINDS = ("RBF", "nearest_neighbor", "linear")

    def __init__(self, reduce_basis_dict, kind="RBF", ctx=None):
        import torch
        if kind not in self.KINDS:
            raise NotImplementedError("Not implemented schur complement computation method.")            # lattice_sim.py:955
        self.ctx = _ctx(ctx)
        self.kind = kind
        basis = np.asarray(reduce_basis_dict["basis_reduced_ortho"], dtype=np.float64)                  # (n*n, k)
        self.length, self.k = basis.shape
        self.n = isqrt(self.length)
        if self.n * self.n != self.length:
            raise ValueError("basis_reduced_ortho does not hold square matrices")
        self.list_elements = np.asarray(reduce_basis_dict["list_elements"], dtype=np.float64)
        if self.list_elements.ndim == 1:
            self.list_elements = self.list_elements[:, None]
        self.alpha_train = np.asarray(reduce_basis_dict["alpha_ortho"], dtype=np.float64).T             # (N, k), :132
        self.d = self.list_elements.shape[1]
        kp = 4 * ((self.k + 3) // 4)
        self._basisP = torch.empty((kp, self.length), dtype=torch.float64, device=self.ctx.device)
        self.ctx.check(self.ctx.lib.lat_basis_prepare(self.ctx.h, _ptr(_dev(self.ctx, basis)), self.length, self.k, self.n,
                                                      _ptr(self._basisP)))
        self._X = _dev(self.ctx, self.list_elements)
        self._A = _dev(self.ctx, self.alpha_train)
        self.rbf = ThinPlateSplineRBF(self.list_elements, self.alpha_train, ctx=self.ctx) if kind == "RBF" else None
        self._tri = None
        if kind == "linear" and self.d > 1:
            # set-up geometry on the host: the Qhull triangulation scipy's LinearNDInterpolator builds (lattice_sim.py:797)
            from scipy.spatial import Delaunay
            tri = Delaunay(self.list_elements)
            self._tri = (_dev(self.ctx, tri.simplices, np.int32), _dev(self.ctx, tri.transform), int(tri.simplices.shape[0]))

    def _queries(self, params):
        xq = np.asarray(params, dtype=np.float64)
        if xq.ndim == 1:
            xq = xq[None, :]
        if xq.shape[1] != self.d:
            raise ValueError(f"queries have {xq.shape[1]} parameters, the surrogate {self.d}")
        return _dev(self.ctx, xq)

    def alphas_device(self, xq):
        import torch
        if self.kind == "RBF":
            return self.rbf.evaluate_device(xq)
        out = torch.empty((xq.shape[0], self.k), dtype=torch.float64, device=self.ctx.device)
        if self._tri is not None:
            simplices, transform, ns = self._tri
            self.ctx.check(self.ctx.lib.lat_alpha_simplex(self.ctx.h, _ptr(simplices), _ptr(transform), ns, self.d, _ptr(self._X),
                                                          self.list_elements.shape[0], _ptr(self._A), self.k, _ptr(xq),
                                                          xq.shape[0], _ptr(out)))
            return out
        self.ctx.check(self.ctx.lib.lat_alpha_lookup(self.ctx.h, 0 if self.kind == "nearest_neighbor" else 1, _ptr(self._X),
                                                     self.list_elements.shape[0], self.d, _ptr(self._A), self.k, _ptr(xq),
                                                     xq.shape[0], _ptr(out)))
        return out

    def expand_device(self, alphas, out=None):
        """(M, k) coefficients -> (M, n, n) matrices on the device."""
        import torch
        M = alphas.shape[0]
        if out is None:
            out = torch.empty((M, self.n, self.n), dtype=torch.float64, device=self.ctx.device)
        self.ctx.check(self.ctx.lib.lat_basis_expand(self.ctx.h, _ptr(self._basisP), self.k, self.length, _ptr(alphas), M,
                                                     alphas.stride(0), _ptr(out)))
        return out

    def schur_batch_device(self, params, out=None):
        return self.expand_device(self.alphas_device(self._queries(params)), out=out)

    def schur_batch(self, params):
        """``get_schur_complement_from_reduced_basis_batch``: (n_queries, n, n) numpy."""
        return self.schur_batch_device(params).cpu().numpy()

    def schur_gradients_device(self, params):
        """(M, d, n, n): dS/d(parameter j) of every query (RBF only, ``_compute_schur_gradients_RBF``)."""
        if self.kind != "RBF":
            raise NotImplementedError("analytic surrogate gradients exist for the RBF surrogate only (lattice_sim.py:893-899)")
        xq = self._queries(params)
        g = self.rbf.gradient_device(xq)                                           # (M, d, k)
        M = g.shape[0]
        return self.expand_device(g.reshape(M * self.d, self.k)).reshape(M, self.d, self.n, self.n)


# ---- drop-ins for the LatticeSim methods (bound by install.patch_reference) --------------------------------------
def _surrogate_of(lattice, ctx=None):
    kind = lattice.type_schur_complement_computation
    cache = getattr(lattice, "_b200_surrogate", None)
    if cache is None or cache.kind != kind:
        cache = SchurSurrogate(lattice.reduce_basis_dict, kind, ctx=ctx)
        lattice._b200_surrogate = cache
        if lattice.shape_schur_complement is None:
            lattice.shape_schur_complement = cache.n                               # lattice_sim.py:966-967
        if kind == "RBF":
            lattice.radial_basis_function = cache.rbf                              # what _define_radial_basis_functions sets
    return cache


def lattice_schur_batch(lattice, geometric_params_list, ctx=None):
    """``LatticeSim.get_schur_complement_from_reduced_basis_batch`` (lattice_sim.py:921-978)."""
    return _surrogate_of(lattice, ctx).schur_batch([list(p) for p in geometric_params_list])


def lattice_schur_single(lattice, geometric_params, ctx=None):
    """``LatticeSim.get_schur_complement_from_reduced_basis`` (lattice_sim.py:980-1018)."""
    return _surrogate_of(lattice, ctx).schur_batch([list(geometric_params)])[0]


def lattice_schur_gradients_rbf(lattice, radii_params, ctx=None):
    """``LatticeSim._compute_schur_gradients_RBF`` (lattice_sim.py:1056-1082): list of dS/dr_j."""
    g = _surrogate_of(lattice, ctx).schur_gradients_device([list(radii_params)])[0].cpu().numpy()
    return [g[j] for j in range(g.shape[0])]


def lattice_define_rbf(lattice, ctx=None):
    """``LatticeSim._define_radial_basis_functions`` (lattice_sim.py:809-812)."""
    lattice.radial_basis_function = ThinPlateSplineRBF(np.array(lattice.reduce_basis_dict["list_elements"]),
                                                       np.array(lattice.alpha_coefficients_greedy), ctx=ctx)
