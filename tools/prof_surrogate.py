"""One launch of each surrogate kernel shape for ncu (see tools/ab_surrogate.py for the timings)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, surrogate
ctx = L.Context()
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
rb = np.load(os.path.join(g, "reduced_basis_BCC_tol_1e-6.npz"))
ref = np.load(os.path.join(g, "surrogate_ref.npz"))
rng = np.random.default_rng(1)
q, _ = np.linalg.qr(rng.standard_normal((84 * 84, 38)))
for rbd, M in ((rb, 216000), ({"basis_reduced_ortho": q, "alpha_ortho": ref["a2"].T, "list_elements": ref["x2"]}, 40000)):
    s = surrogate.SchurSurrogate(rbd, "RBF", ctx=ctx)
    xq = surrogate._dev(ctx, rng.uniform(s.list_elements.min(0), s.list_elements.max(0), (M, s.d)))
    out = torch.empty((M, s.n, s.n), dtype=torch.float64, device=ctx.device)
    for _ in range(2):
        s.expand_device(s.alphas_device(xq), out=out)
    torch.cuda.synchronize()
