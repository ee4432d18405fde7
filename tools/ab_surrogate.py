"""Throughput of the surrogate Schur evaluation (csrc/lattice_surrogate.cu): RBF alphas + DMMA basis expansion.

    python tools/ab_surrogate.py            # BASELINE config 4 size (216 000 cells) from the reference's BCC basis, and a
                                            # 2-parameter / 38-vector / 84x84 case shaped like reduced_basis_BCC_Hybrid4
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, surrogate  # noqa: E402

PEAK = 6554.6


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def case(ctx, name, rb, M, d_grad=False):
    s = surrogate.SchurSurrogate(rb, "RBF", ctx=ctx)
    rng = np.random.default_rng(0)
    lo, hi = s.list_elements.min(0), s.list_elements.max(0)
    xq = surrogate._dev(ctx, rng.uniform(lo, hi, (M, s.d)))
    out = torch.empty((M, s.n, s.n), dtype=torch.float64, device=ctx.device)
    al = s.alphas_device(xq)
    t_al = timed(lambda: s.alphas_device(xq))
    t_ex = timed(lambda: s.expand_device(al, out=out))
    by = M * s.length * 8
    fl = 2.0 * M * s.length * (4 * ((s.k + 3) // 4))
    print(f"{name} [LAT_EXPAND_MT={os.environ.get('LAT_EXPAND_MT', 'default')}]: M={M} n={s.n} k={s.k} d={s.d} centres={s.list_elements.shape[0]}")
    print(f"   RBF alphas (k_tps_eval)        : {t_al:8.3f} ms")
    print(f"   basis @ alphas (k_basis_expand): {t_ex:8.3f} ms   S written at {by / t_ex / 1e6:7.0f} GB/s = {by / t_ex / 1e6 / PEAK:.2f} of HBM;"
          f"  DMMA {fl / t_ex / 1e9:6.2f} TFLOP/s;  {M / t_ex / 1e3:7.2f} M cells/s")
    if d_grad:
        g = s.rbf.gradient_device(xq).reshape(M * s.d, s.k)
        og = torch.empty((M * s.d, s.n, s.n), dtype=torch.float64, device=ctx.device)
        t_g = timed(lambda: s.rbf.gradient_device(xq))
        t_ge = timed(lambda: s.expand_device(g, out=og))
        print(f"   gradient alphas                : {t_g:8.3f} ms;  dS = basis @ dalpha: {t_ge:8.3f} ms "
              f"({M * s.d * s.length * 8 / t_ge / 1e6:7.0f} GB/s)")


def main():
    ctx = L.Context()
    g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    rb = np.load(os.path.join(g, "reduced_basis_BCC_tol_1e-6.npz"))
    case(ctx, "config 4 (BCC 60^3 cells, reference basis tol 1e-6)", rb, 216000, d_grad=True)
    ref = np.load(os.path.join(g, "surrogate_ref.npz"))
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.standard_normal((84 * 84, 38)))
    case(ctx, "2-parameter hybrid shape (84x84, 38 vectors, 100 centres of the reference's BCC_Hybrid4 set)",
         {"basis_reduced_ortho": q, "alpha_ortho": ref["a2"].T, "list_elements": ref["x2"]}, 40000)


if __name__ == "__main__":
    main()
