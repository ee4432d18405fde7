"""Drop-in for ``pyLatticeSim.conjugate_gradient_solver.conjugate_gradient_solver``
(conjugate_gradient_solver.py:15-122) for operators that live on the GPU as BSR(6x6) matrices.

Same signature, same ``(x, info)`` return, same iteration (x0 = 0, alpha clamp, periodic restart, the two
stop tests, info 0/1/2) -- executed by ``lat_pcg_bsr(reference_semantics=1)``.  There is no CPU fallback:
an operator that is not a :class:`BsrOperator` is a ``TypeError``.
"""
from __future__ import annotations

import numpy as np

from . import lib as L


class BsrOperator:
    """Assembled operator on the device; ``A @ v`` works for numpy vectors (host round trip) and torch tensors."""

    def __init__(self, ctx: L.Context, rowptr, colidx, vals):
        self.ctx, self.rowptr, self.colidx, self.vals = ctx, rowptr, colidx, vals
        n = 6 * (int(rowptr.numel()) - 1)
        self.shape = (n, n)

    def __matmul__(self, v):
        import torch
        if torch.is_tensor(v):
            return self.ctx.spmv(self.rowptr, self.colidx, self.vals, v)
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(self.ctx.device)
        return self.ctx.spmv(self.rowptr, self.colidx, self.vals, t).cpu().numpy()

    matvec = __matmul__


class Jacobi:
    """Marker for ``M``: diagonal preconditioner built on the device from the operator."""
    kind = L.PC_JACOBI


class BlockJacobi:
    """Marker for ``M``: 6x6 block-Jacobi preconditioner built on the device from the operator."""
    kind = L.PC_BLOCK6


def conjugate_gradient_solver(A_operator, b, M=None, maxiter=100, tol=1e-5, mintol=1e-5, restart_every=1000,
                              alpha_max=0.1, callback=None):
    """Solve A x = b with the reference's PCG on the GPU.  ``M``: None, :class:`Jacobi` or
    :class:`BlockJacobi` (class or instance).  ``callback`` is invoked once, with the final iterate
    (the reference calls it every iteration, :85-86; the device loop does not return to the host per iteration)."""
    import torch
    if not isinstance(A_operator, BsrOperator):
        raise TypeError("conjugate_gradient_solver (B200): A_operator must be a BsrOperator; there is no CPU fallback")
    kind = L.PC_NONE if M is None else getattr(M, "kind", None)
    if kind is None:
        raise TypeError("M must be None, Jacobi or BlockJacobi")
    ctx = A_operator.ctx
    is_t = torch.is_tensor(b)
    bt = b if is_t else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float64)).to(ctx.device)
    x, info = ctx.pcg(A_operator.rowptr, A_operator.colidx, A_operator.vals, bt, tol=float(tol), maxiter=int(maxiter),
                      precond=kind, reference_semantics=True, mintol=float(mintol), alpha_max=float(alpha_max),
                      restart_every=int(restart_every))
    out = x if is_t else x.cpu().numpy()
    if callback is not None:
        callback(out)
    return out, int(info["info"])
