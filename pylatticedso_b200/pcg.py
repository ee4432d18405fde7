"""Drop-in for ``pyLatticeSim.conjugate_gradient_solver.conjugate_gradient_solver``
(conjugate_gradient_solver.py:15-122) for operators that live on the GPU as BSR(6x6) matrices.

Same signature, same ``(x, info)`` return, same iteration (x0 = 0, alpha clamp, periodic restart, the two
stop tests, info 0/1/2) -- executed by ``lat_pcg_bsr(reference_semantics=1)``.

Two kinds of ``A_operator`` are accepted:

* a :class:`BsrOperator` (an assembled matrix already on the device);
* the operators the reference itself builds: a ``scipy.sparse.linalg.LinearOperator`` whose ``matvec`` is
  ``LatticeSim.calculate_reaction_force_global`` (``solve_DDM``, lattice_sim.py:1148-1160) or a closure around it
  (``LatticeOpti._solve_adjoint_vector``, lattice_opti.py:1636-1645).  The bound lattice is recognised, its interface
  operator ``sum_c B_c S_c B_c^T`` is assembled on the device from ``cell.schur_complement`` and the iteration runs
  there on the free interface DOFs -- the Python loop over all cells per iteration never runs.

There is no CPU fallback: any other operator is a ``TypeError``.
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


def lattice_of_operator(A_operator):
    """The LatticeSim behind a reference-built LinearOperator, or None."""
    fn = getattr(A_operator, "_CustomLinearOperator__matvec_impl", None)
    if fn is None:
        fn = getattr(A_operator, "matvec", None)
    owner = getattr(fn, "__self__", None)
    if owner is not None and hasattr(owner, "calculate_reaction_force_global") and hasattr(owner, "cells"):
        return owner
    for cell in (getattr(fn, "__closure__", None) or ()):
        try:
            obj = cell.cell_contents
        except ValueError:
            continue
        if hasattr(obj, "calculate_reaction_force_global") and hasattr(obj, "cells"):
            return obj
    return None


class DdmOperator:
    """Interface operator of a decomposed lattice on the device, in the reference's free-DOF numbering.
    ``A @ v`` (numpy) = ``LatticeSim.calculate_reaction_force_global(v)`` (lattice_sim.py:1180-1200)."""

    def __init__(self, lattice, ctx=None):
        from . import ddm
        if getattr(lattice, "free_DOF", None) is None:
            lattice.define_free_DOF()
        lattice.set_global_free_DOF_index()
        self.lattice = lattice
        self.prob, self.pts, self.fixed, _, _ = ddm.interface_from_lattice(lattice, ctx)
        self.ctx = self.prob.ctx
        self.free = ddm.free_dof_map(lattice, self.pts)
        n = int(self.free.shape[0])
        self.shape = (n, n)

    def __matmul__(self, v):
        import torch
        full = np.zeros(self.fixed.shape[0])
        full[self.free] = np.asarray(v, dtype=np.float64)
        t = torch.from_numpy(full).to(self.ctx.device)
        y = self.ctx.spmv(self.prob.rowptr, self.prob.colidx, self.prob.vals, t).cpu().numpy()
        return y[self.free]

    matvec = __matmul__


def _solve_ddm_operator(op: DdmOperator, b, kind, maxiter, tol, mintol, restart_every, alpha_max):
    """Reference-semantics PCG on the free interface DOFs: the system is embedded in the 6-DOF-per-node BSR matrix
    with the constrained rows / columns replaced by the identity and a zero right-hand side there, which leaves the
    iterates on the free DOFs (and every norm the stop tests use) unchanged."""
    import torch
    ctx = op.ctx
    dev = ctx.device
    n_full = op.fixed.shape[0]
    f = np.zeros(n_full)
    f[op.free] = np.asarray(b, dtype=np.float64)
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
    vbc, rhs = ctx.apply_dirichlet(op.prob.rowptr, op.prob.colidx, op.prob.vals, t(op.fixed, np.uint8),
                                   t(np.zeros(n_full), np.float64), t(f, np.float64))
    x, info = ctx.pcg(op.prob.rowptr, op.prob.colidx, vbc, rhs, tol=float(tol), maxiter=int(maxiter), precond=kind,
                      reference_semantics=True, mintol=float(mintol), alpha_max=float(alpha_max),
                      restart_every=int(restart_every))
    return x.cpu().numpy()[op.free], info


def conjugate_gradient_solver(A_operator, b, M=None, maxiter=100, tol=1e-5, mintol=1e-5, restart_every=1000,
                              alpha_max=0.1, callback=None):
    """Solve A x = b with the reference's PCG on the GPU.

    ``A_operator``: :class:`BsrOperator`, :class:`DdmOperator` or a reference-built LinearOperator (module docstring).
    ``M``: None (no preconditioner), :class:`Jacobi` / :class:`BlockJacobi` (class or instance); any other object --
    the reference passes its SuperLU-based ``LinearOperator`` (lattice_sim.py:1333-1415) -- selects the device's 6x6
    block-Jacobi preconditioner, because a host-side factorisation cannot be applied inside the device loop (iteration
    counts then differ from the reference's, the stopping rules do not).
    ``callback`` is invoked once, with the final iterate (the reference calls it every iteration, :85-86; the device
    loop does not return to the host per iteration)."""
    import torch
    lattice = None if isinstance(A_operator, (BsrOperator, DdmOperator)) else lattice_of_operator(A_operator)
    if lattice is not None:
        A_operator = DdmOperator(lattice)
    if not isinstance(A_operator, (BsrOperator, DdmOperator)):
        raise TypeError("conjugate_gradient_solver (B200): A_operator must be a BsrOperator, a DdmOperator or the "
                        "LinearOperator LatticeSim builds around calculate_reaction_force_global; there is no CPU fallback")
    if M is None:
        kind = L.PC_NONE
    else:
        kind = getattr(M, "kind", L.PC_BLOCK6)
    if isinstance(A_operator, DdmOperator):
        out, info = _solve_ddm_operator(A_operator, b, kind, maxiter, tol, mintol, restart_every, alpha_max)
        if callback is not None:
            callback(out)
        return out, int(info["info"])
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
