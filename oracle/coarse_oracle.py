"""CPU restatement of the two-level preconditioner (pylatticedso_b200/csrc/coarse.cuh) -- TEST INFRASTRUCTURE ONLY.

The reference has no counterpart of this preconditioner: it passes SuperLU's factorisation of the whole interface
matrix to its PCG (lattice_sim.py:1333-1415).  What parity means here: the SOLUTION of the preconditioned solve equals
the reference solution of the same linear system (oracle/lattice_oracle.py: solve_static, pinned by the reference's
fixtures) to the solver tolerance, whatever the preconditioner; this file restates the preconditioner itself in
plain numpy / scipy.sparse so that the device's coarse matrix, one application of the coarse correction and the
iteration counts can be checked against an independent implementation.  Parity of the preconditioner as such is
therefore "unpinned" (no reference vectors exist); parity of the solve is pinned through lattice_oracle.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def box_aggregates(xyz, target):
    """Same box rule as pylatticedso_b200.coarse.box_aggregates (numpy)."""
    lo, hi = xyz.min(0), xyz.max(0)
    ext = np.maximum(hi - lo, 1e-300)
    live = ext > 1e-9 * ext.max()
    h = (np.prod(ext[live]) / target) ** (1.0 / max(1, int(live.sum())))
    nb = np.where(live, np.maximum(1, np.round(ext / h)), 1).astype(np.int64)
    ijk = np.clip(np.floor((xyz - lo) / ext * nb).astype(np.int64), 0, nb - 1)
    flat = (ijk[:, 0] * nb[1] + ijk[:, 1]) * nb[2] + ijk[:, 2]
    uniq, inv = np.unique(flat, return_inverse=True)
    return inv, len(uniq)


def rigid_body_modes(xyz, agg, n_agg, fixed=None):
    """Z (6 n_nodes x 6 n_agg, CSR): per node the block M_i [I, -[d_i]x; 0, I] in the columns of its aggregate."""
    n = xyz.shape[0]
    cnt = np.bincount(agg, minlength=n_agg).astype(float)
    cen = np.stack([np.bincount(agg, xyz[:, k], minlength=n_agg) / np.maximum(cnt, 1.0) for k in range(3)], 1)
    d = (xyz - cen[agg]).astype(np.float32).astype(np.float64)   # the device keeps the lever arms in FP32 (CoarseNode)
    nodes = np.arange(n)
    rows, cols, vals = [], [], []
    for k in range(6):                                   # identity blocks
        rows.append(6 * nodes + k); cols.append(6 * agg + k); vals.append(np.ones(n))
    # u = w x d:  u_x = w_y d_z - w_z d_y,  u_y = w_z d_x - w_x d_z,  u_z = w_x d_y - w_y d_x
    for ui, wk, comp, sgn in ((0, 1, 2, 1.0), (0, 2, 1, -1.0), (1, 2, 0, 1.0), (1, 0, 2, -1.0), (2, 0, 1, 1.0), (2, 1, 0, -1.0)):
        rows.append(6 * nodes + ui); cols.append(6 * agg + 3 + wk); vals.append(sgn * d[:, comp])
    Z = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * n, 6 * n_agg))
    if fixed is not None:
        Z = sp.diags((~np.asarray(fixed, bool)).astype(float)) @ Z
    return Z.tocsr()


def coarse_matrix(K, Z):
    return np.asarray((Z.T @ (K @ Z)).todense())


def coarse_pinv(E, cut=1e-11):
    E = 0.5 * (E + E.T)
    w, V = np.linalg.eigh(E)
    keep = w > cut * w.max()
    return (V[:, keep] / w[keep]) @ V[:, keep].T


def block_jacobi_inverse(K, n_nodes):
    B = sp.bsr_matrix(K, blocksize=(6, 6))
    B.sort_indices()
    D = np.zeros((n_nodes, 6, 6))
    row = np.repeat(np.arange(n_nodes), np.diff(B.indptr))
    diag = B.indices == row
    D[row[diag]] = B.data[diag]
    return np.linalg.inv(D)


def two_level_pcg(K, b, Dinv, Z=None, Einv=None, tol=1e-8, maxiter=100000):
    """Textbook PCG (r0 = b, stop |r| <= tol |b|) with M^-1 = D^-1 (+ Z Einv Z^T).  Returns (x, iterations)."""
    n_nodes = Dinv.shape[0]

    def prec(r):
        u = np.einsum("nij,nj->ni", Dinv, r.reshape(n_nodes, 6)).ravel()
        if Z is not None:
            u = u + Z @ (Einv @ (Z.T @ r))
        return u

    x = np.zeros_like(b)
    r = b.copy()
    u = prec(r)
    p = u.copy()
    gamma = r @ u
    bb = b @ b
    it = 0
    while it < maxiter and r @ r > tol * tol * bb:
        w = K @ p
        alpha = gamma / (p @ w)
        x += alpha * p
        r -= alpha * w
        u = prec(r)
        g2 = r @ u
        p = u + (g2 / gamma) * p
        gamma = g2
        it += 1
    return x, it
