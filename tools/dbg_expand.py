"""Locate mismatches of lat_basis_expand against numpy (debug aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, surrogate
from oracle import surrogate_oracle as so
ctx = L.Context()
for n, k, M in ((42, 70, 65), (48, 70, 65), (42, 70, 64), (42, 66, 8), (42, 130, 40)):
    rng = np.random.default_rng(n * 1000 + k)
    basis = rng.standard_normal((n * n, k)); alphas = rng.standard_normal((M, k))
    s = surrogate.SchurSurrogate({"basis_reduced_ortho": basis, "alpha_ortho": np.zeros((k, 2)), "list_elements": np.zeros((2, 1))}, "nearest_neighbor", ctx=ctx)
    got = s.expand_device(surrogate._dev(ctx, alphas)).cpu().numpy().reshape(M, -1)
    want = so.schur_from_alphas(basis, alphas, n).reshape(M, -1)
    bad = np.argwhere(np.abs(got - want) > 1e-9)
    print(n, k, M, "bad:", len(bad), "rows", np.unique(bad[:, 0])[:70], "cols", np.unique(bad[:, 1])[:40], "...", np.unique(bad[:, 1])[-10:] if len(bad) else "")
    if len(bad):
        # which k-range explains the difference?
        d = got - want
        r, c = bad[0]
        a, b = c // n, c % n
        contrib = alphas[r] * basis[a + n * b]
        print("   first bad", r, c, "diff", d[r, c], "sum k<64", contrib[:64].sum(), "sum k>=64", contrib[64:].sum(), "got", got[r, c], "want", want[r, c])
