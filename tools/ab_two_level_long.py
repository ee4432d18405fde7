"""The weak-scaled headline domain on ONE GPU: BCC (20 N) x 20 x 20, m = 2, block-Jacobi against two-level (does the coarse
space remove the iteration growth with the domain length?).   python tools/ab_two_level_long.py [N ...]"""
import sys

import torch

sys.path.insert(0, ".")
from pylatticedso_b200 import lib as L
from pylatticedso_b200 import mesh as M
from pylatticedso_b200.fem import BeamFEM

ctx = L.Context()
for n in [int(a) for a in sys.argv[1:]] or [1, 4, 8]:
    m = M.mesh_from_synthetic(M.synthetic_lattice("BCC", (20 * n, 20, 20), [0.05]), 2)
    fixed, g, f = M.compression_bc(m)
    fem = BeamFEM(m, 1013.0, 0.3, ctx=ctx)
    for solve, name in ((fem.solve_matrix_free, "matrix-free"),):
        u0, _, i0 = solve(fixed, g, f, tol=1e-8, want_reactions=False)
        tl = fem.two_level(fixed)
        u1, _, i1 = solve(fixed, g, f, tol=1e-8, want_reactions=False, two_level=tl)
        print(f"BCC {20 * n}x20x20 m=2 ({m.n_dof} DOF) {name}: block-Jacobi {i0['iters']} it {i0['solve_ms']:.1f} ms | two-level "
              f"({tl.n_agg} aggregates) {i1['iters']} it {i1['solve_ms']:.1f} ms | |du|/|u| {float((u1 - u0).abs().max() / u0.abs().max()):.1e}", flush=True)
        del tl
    del fem
    torch.cuda.empty_cache()
