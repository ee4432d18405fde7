"""CPU restatement of the reference's reduced-basis / surrogate Schur pipeline.

TEST INFRASTRUCTURE ONLY: imported by tests/, never by the package.  Plain numpy, no BLAS/LAPACK driver calls
beyond ``np.linalg.solve`` / ``np.linalg.lstsq``.  Pinned by (tests/test_oracle_surrogate.py)

* the reference's stored reduced bases ``data/outputs/schur_complement/reduced_basis/reduced_basis_{BCC_tol_1e-3,
  BCC_tol_1e-6,Hybrid1_tol_1e-6,Hybrid4_tol_1e-6}.npz`` (copied to tests/golden/), which the greedy below reproduces
  from the reference's stored Schur datasets (tests/golden/schur_*.npz), and
* outputs of the reference's own ``reduce_basis_greedy`` / ``ThinPlateSplineRBF`` /
  ``LatticeSim.get_schur_complement_from_reduced_basis_batch`` / ``_compute_schur_gradients_RBF`` /
  ``evaluate_alphas_linear_surrogate`` run in the build container and frozen by
  tests/golden/make_golden_surrogate.py (tests/golden/surrogate_ref.npz).
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# greedy reduced basis -- greedy_algorithm.py:35-155
# ---------------------------------------------------------------------------------------------------------------
def greedy_reduced_basis(schur_dict: dict, tol_greedy: float):
    """Returns the reference's 7-tuple (mainelem, reducedcoef, projfieldpp, basis_reduced_ortho, alpha_ortho,
    matP_sorted, norm_mainelem_sorted)."""
    keys = sorted(schur_dict.keys())                                     # :87
    mats = np.array([schur_dict[k] for k in keys])                       # :89
    n_snap = len(keys)
    norms = np.ones(n_snap)
    fields = []
    for i in range(n_snap):                                              # :98-102 vec_F(S) / |vec_F(S)|
        v = np.ravel(mats[i], order="F")
        norms[i] = np.sqrt(np.sum(v * v))
        fields.append(v / norms[i])
    D = np.stack(fields).T.copy()                                        # (len, n_snap)
    atol = tol_greedy * np.max(np.sum(np.abs(D), axis=0))                # :105: norm(D.T, inf) = max column 1-norm
    basis, coefs, main = [], [], []
    count, cvg = 0, False
    while (not cvg) and count < n_snap:                                  # :112
        count += 1
        s_i = int(np.argmax(np.max(np.abs(D), axis=0)))                  # :114-115 column inf-norms, first maximum
        newvec = D[:, s_i] / np.sqrt(np.sum(D[:, s_i] ** 2))             # :117
        newcoef = D.T @ newvec                                           # :118 dgemv(trans)
        D -= np.outer(newvec, newcoef)                                   # :119 dger
        cvg = bool(np.max(np.sum(np.abs(D), axis=0)) < atol)             # :121
        basis.append(newvec.copy()); coefs.append(newcoef.copy()); main.append(s_i)
    main = np.array(main)
    coef = np.stack(coefs)                                               # (k, n_snap)
    matP = np.triu(coef[:, main])                                        # :128
    coef = np.linalg.solve(matP, coef)                                   # :129 dtrtrs, upper triangular
    coef = coef * np.outer(1.0 / norms[main], norms)                     # :130
    vsort = np.argsort(main)
    B = np.column_stack(basis)
    alpha = np.zeros((B.shape[1], n_snap))
    for s in range(n_snap):                                              # :135-138
        alpha[:, s] = np.linalg.lstsq(B, np.ravel(mats[s], order="F"), rcond=None)[0]
    return (main[vsort], coef[vsort, :], [mats[i] for i in main[vsort]], B, alpha,
            matP[np.ix_(vsort, vsort)], norms[main[vsort]])


def project_to_basis(schur_dict: dict, basis: np.ndarray):
    """greedy_algorithm.py:233-266 (note: C-order ravel there, F-order in the greedy; S is symmetric)."""
    return {k: np.linalg.lstsq(basis, np.ravel(S, order="C").astype(float), rcond=None)[0] for k, S in schur_dict.items()}


# ---------------------------------------------------------------------------------------------------------------
# thin-plate-spline RBF -- utils_rbf.py:13-144
# ---------------------------------------------------------------------------------------------------------------
def _phi(r):
    out = np.zeros_like(r)
    m = r > 0
    out[m] = r[m] ** 2 * np.log(r[m])                                    # :66-71
    return out


def tps_fit(x_train, y_train, reg=0.0):
    X = np.asarray(x_train, float)
    Y = np.asarray(y_train, float)
    if Y.ndim == 1:
        Y = Y[:, None]
    N, d = X.shape
    Phi = _phi(np.linalg.norm(X[:, None, :] - X[None, :, :], axis=2))
    if reg > 0.0:
        Phi = Phi + reg * np.eye(N)
    P = np.hstack([np.ones((N, 1)), X])
    A = np.block([[Phi, P], [P.T, np.zeros((d + 1, d + 1))]])            # :50-53
    sol = np.linalg.solve(A, np.vstack([Y, np.zeros((d + 1, Y.shape[1]))]))
    return sol                                                           # rows [0, N): W, rows [N, N+d+1): CP


def tps_evaluate(x_train, wcp, xq):
    X = np.asarray(x_train, float)
    Q = np.atleast_2d(np.asarray(xq, float))
    N = X.shape[0]
    r = np.linalg.norm(Q[:, None, :] - X[None, :, :], axis=2)
    return _phi(r) @ wcp[:N] + np.hstack([np.ones((Q.shape[0], 1)), Q]) @ wcp[N:]      # :101-103


def tps_gradient(x_train, wcp, xq):
    X = np.asarray(x_train, float)
    Q = np.atleast_2d(np.asarray(xq, float))
    N = X.shape[0]
    Dv = Q[:, None, :] - X[None, :, :]
    r = np.linalg.norm(Dv, axis=2)
    fac = np.zeros_like(r)
    m = r > 0
    fac[m] = 2.0 * np.log(r[m]) + 1.0                                    # :126-128
    G = np.einsum("qnd,nk->qdk", fac[:, :, None] * Dv, wcp[:N])          # :135
    return G + wcp[N + 1:][None, :, :]                                   # :138-139


# ---------------------------------------------------------------------------------------------------------------
# alpha look-ups and reconstruction -- lattice_sim.py:755-807, 921-978, 1056-1082
# ---------------------------------------------------------------------------------------------------------------
def alphas_nearest(list_elements, alpha_train, xq):
    X = np.asarray(list_elements, float)
    Q = np.atleast_2d(np.asarray(xq, float))
    idx = np.argmin(np.linalg.norm(Q[:, None, :] - X[None, :, :], axis=2), axis=1)     # 1-NN (:941)
    return np.asarray(alpha_train)[idx]


def alphas_linear_1d(list_elements, alpha_train, xq):
    """1-D branch of evaluate_alphas_linear_surrogate (:781-792): np.interp per coefficient, clamped outside."""
    x = np.asarray(list_elements, float).ravel()
    order = np.argsort(x)
    xs, As = x[order], np.asarray(alpha_train, float)[order]
    q = np.asarray(xq, float).ravel()
    return np.stack([np.interp(q, xs, As[:, j]) for j in range(As.shape[1])], axis=1)


def alphas_linear_nd(list_elements, alpha_train, xq):
    """N-parameter branch of evaluate_alphas_linear_surrogate (:794-807): scipy's LinearNDInterpolator inside the convex
    hull of the centres, NearestNDInterpolator where it returns NaN."""
    from scipy.interpolate import LinearNDInterpolator, NearestNDInterpolator
    X, A = np.asarray(list_elements, float), np.asarray(alpha_train, float)
    lin, nn = LinearNDInterpolator(X, A), NearestNDInterpolator(X, A)
    Q = np.atleast_2d(np.asarray(xq, float))
    y = lin(Q)
    bad = np.isnan(y).any(axis=1)
    if bad.any():
        y[bad] = nn(Q[bad])
    return y


def schur_from_alphas(basis, alphas, n):
    """(n_q, n, n) from basis (n*n, k) and alphas (n_q, k): S_flat = basis @ alphas.T, each column reshaped in
    Fortran order (:961-976)."""
    flat = np.asarray(basis) @ np.asarray(alphas).T
    return np.stack([flat[:, j].reshape((n, n), order="F") for j in range(flat.shape[1])])
