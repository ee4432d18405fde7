"""k_schur_dense (row A7) at the reference's own mesh density: BCC cell, 18 elements per strut (gmsh rule h = 0.05 cell size:
137 interior nodes, 870 DOF, 48 boundary DOF) and coarser cells for comparison."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L
from pylatticedso_b200.schur import synthetic_cell_batch
ctx = L.Context()
rng = np.random.default_rng(44)
for m_, nc in ((1, 216000), (3, 50000), (18, 4000), (18, 216000)):
    if m_ == 18 and nc == 216000:
        # full config 4 at the reference density: only the pre-pass path (the dense path needs 3 s)
        radii = 0.02 + 0.06 * rng.random(nc)
        batch, bnd = synthetic_cell_batch(ctx, "BCC", radii, m_, 1013.0, 0.3)
        S = batch.schur(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S = batch.schur(); e1.record(); torch.cuda.synchronize()
        print(f"lat_schur_batch BCC m=18 nI= 822 cells=216000 strut pre-pass + joints : {e0.elapsed_time(e1):9.2f} ms  (BASELINE config 4 at the reference mesh density)", flush=True)
        continue
    radii = 0.02 + 0.06 * rng.random(nc)
    batch, bnd = synthetic_cell_batch(ctx, "BCC", radii, m_, 1013.0, 0.3)
    nI = 6 * (batch.xyz.shape[1] - 8)
    res = {}
    for name, kw in (("dense (all interior nodes)", dict(use_chains=False)), ("strut pre-pass + joints ", dict(use_chains=True))):
        if name.startswith("strut") and batch.chains is None:
            continue
        S = batch.schur(**kw); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S = batch.schur(**kw); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res[name] = S
        print(f"lat_schur_batch BCC m={m_:2d} nI={nI:4d} cells={nc:6d} {name}: {ms:9.2f} ms  {nc/ms/1e3:8.3f} M cells/s  -> 216000 cells (BCC 60^3) in {216000/nc*ms/1e3:7.3f} s", flush=True)
    if len(res) == 2:
        a, b = res.values()
        print(f"      max relative difference between the two paths: {float((a - b).abs().max() / a.abs().max()):.1e}", flush=True)
    del res, batch
