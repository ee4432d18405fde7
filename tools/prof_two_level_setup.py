"""Where the set-up time of the two-level preconditioner goes (Octet n^3, default 100):  python tools/prof_two_level_setup.py [n]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pylatticedso_b200 import coarse, lib as L, mesh as M
from pylatticedso_b200.fem import BeamFEM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = L.Context()
m = M.mesh_from_synthetic(M.synthetic_lattice("Octet", (n, n, n), [0.03]), 1)
fixed, g, f = M.compression_bc(m)
fem = BeamFEM(m, 1013.0, 0.3, ctx=ctx)
fem.build_pattern()
fx = torch.from_numpy(fixed.astype(np.uint8)).to(ctx.device)


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, 1e3 * (time.perf_counter() - t0)


for rep in range(2):
    (agg, n_agg), t_box = timed(lambda: coarse.box_aggregates(fem.x, fem.y, fem.z, coarse.default_aggregates(m.n_nodes)))

    def tables():
        order = torch.argsort(agg, stable=True)
        counts = torch.bincount(agg, minlength=n_agg)
        ptr = torch.zeros(n_agg + 1, dtype=torch.int32, device=ctx.device)
        ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        return ptr, order.to(torch.int32).contiguous(), agg.to(torch.int32).contiguous()
    (ptr, nodes, node_agg), t_tab = timed(tables)
    _, t_setup = timed(lambda: ctx.coarse_setup(fem.x, fem.y, fem.z, node_agg, ptr, nodes, fx, None))
    vals, t_asm = timed(lambda: ctx.assemble_bsr(fem.x, fem.y, fem.z, fem.en0, fem.en1, fem.rad, fem.n_nodes, fem.nnzb, fem.young, fem.nu, fem.kappa))
    E, t_gal = timed(lambda: ctx.coarse_galerkin(fem.rowptr, fem.colidx, vals, n_agg))
    Einv, t_inv = timed(lambda: coarse.invert_coarse(E))
    print(f"Octet {n}^3 ({m.n_dof} DOF), {n_agg} aggregates (n_c = {6 * n_agg}), pass {rep}: boxes {t_box:.1f} ms | sort + tables {t_tab:.1f} | "
          f"lat_coarse_setup {t_setup:.1f} | temporary assembly {t_asm:.1f} | Galerkin {t_gal:.1f} | dense inverse {t_inv:.1f} | "
          f"sum {t_box + t_tab + t_setup + t_asm + t_gal + t_inv:.1f} ms", flush=True)
    del vals, E, Einv
