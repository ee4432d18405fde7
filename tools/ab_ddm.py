"""k_ddm_matvec (row A8) timing: 216 000 BCC cells, nb = 48 (4.0 GB of Schur matrices), and Octet-sized nb = 84."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L
ctx = L.Context(); dev = ctx.device
rng = np.random.default_rng(44)
for nb, nc in ((48, 216000), (84, 64000)):
    S = torch.randn((nc, nb, nb), dtype=torch.float64, device=dev)
    gidx = torch.from_numpy(rng.integers(-1, 2_000_000, size=(nc, nb)).astype(np.int32)).to(dev)
    xx = torch.randn(2_000_000, dtype=torch.float64, device=dev); yy = torch.empty_like(xx)
    for _ in range(3): ctx.ddm_matvec(S, gidx, xx, out=yy)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): ctx.ddm_matvec(S, gidx, xx, out=yy)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    by = nc * (nb * nb * 8 + 12 * nb)
    # check against torch (non-symmetric S on purpose: the kernel must apply S, not S^T)
    g = gidx.long(); xg = torch.where(g >= 0, xx[g.clamp(min=0)], torch.zeros((), dtype=torch.float64, device=dev))
    yc = torch.bmm(S, xg.unsqueeze(2)).squeeze(2)
    ref = torch.zeros_like(xx); m = g >= 0
    ref.index_add_(0, g[m], yc[m])
    err = float((yy - ref).abs().max() / ref.abs().max())
    print(f"k_ddm_matvec nb={nb} cells={nc}: {ms*1e3:8.1f} us  {nc/ms/1e3:7.1f} M cells/s  {by/ms/1e6:6.0f} GB/s = {by/ms/1e6/6554.6:.2f} of HBM peak   max rel err vs torch.bmm {err:.1e}", flush=True)
    del S
