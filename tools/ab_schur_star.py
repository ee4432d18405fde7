"""BASELINE config 4 (BCC 60^3 = 216 000 cells, per-cell radii): Schur complements and analytic sensitivities through
the half-warp star kernel (lat_schur_batch_struts) vs the round-1 paths (strut pre-pass + k_schur_dense<SUPER>; dense)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L
from pylatticedso_b200.schur import synthetic_cell_batch
ctx = L.Context()
rng = np.random.default_rng(44)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        a, b = ev(), ev(); a.record(); out = fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best, out
nc = 216000
radii = 0.02 + 0.06 * rng.random(nc)
for m_ in (1, 18):
    batch, bnd = synthetic_cell_batch(ctx, "BCC", radii, m_, 1013.0, 0.3, with_gradients=True)
    out_bytes = nc * 48 * 48 * 8
    ms, S = timed(lambda: batch.schur())
    print(f"BCC m={m_:2d} {nc} cells  star kernel, S only        : {ms:8.3f} ms  {nc/ms/1e3:7.2f} M cells/s  S written at {out_bytes/ms/1e6:6.0f} GB/s", flush=True)
    ms2, S2 = timed(lambda: ctx.schur_batch_chains(batch.xyz, batch.len0, batch.len1, batch.rad, batch.chains, batch.n_bnd_nodes, 1013.0, 0.3, 0.9), n=2)
    print(f"BCC m={m_:2d} {nc} cells  round-1 pre-pass + k_schur_dense<SUPER>: {ms2:8.3f} ms   max rel diff {float((S - S2).abs().max() / S.abs().max()):.1e}", flush=True)
    del S2
    torch.cuda.empty_cache()
    msg, (Sg, dS) = timed(lambda: batch.schur(with_gradients=True), n=2)
    print(f"BCC m={m_:2d} {nc} cells  star kernel, S + dS/dr      : {msg:8.3f} ms  {nc/msg/1e3:7.2f} M cells/s  S+dS written at {2*out_bytes/msg/1e6:6.0f} GB/s", flush=True)
    # spot check of dS against the dense route on a sub-batch
    sub, _ = synthetic_cell_batch(ctx, "BCC", radii[:2000], m_, 1013.0, 0.3, with_gradients=True)
    msd, (Sd, dSd) = timed(lambda: sub.schur(with_gradients=True, use_chains=False), n=1)
    print(f"      dense route (2000 cells): {msd:8.3f} ms -> {nc/2000*msd/1e3:7.3f} s for {nc};  dS rel diff {float((dS[:2000] - dSd).abs().max() / dSd.abs().max()):.1e}", flush=True)
    del batch, S, Sg, dS, sub, Sd, dSd
    torch.cuda.empty_cache()
