"""Interface assembly (A9) at BASELINE config 3 (BCC 60^3, 216 000 cells): InterfaceProblem set-up (device pair generation +
pattern + plan) and the plan-driven gather assembly against the FP64-RED scatter assembly.

    python tools/ab_interface_assembly.py
"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from pylatticedso_b200 import lib as L, ddm, mesh as M
from pylatticedso_b200.schur import bcc_cell_order_nodes, synthetic_cell_batch
ctx = L.Context()
n = 60
rng = np.random.default_rng(44)
radii = 0.02 + 0.06 * rng.random(n ** 3)
batch, _ = synthetic_cell_batch(ctx, "BCC", radii, 1, 1013.0, 0.3)
S = batch.schur()
unit = M.synthetic_lattice("BCC", (1, 1, 1), [1.0])
order = bcc_cell_order_nodes(unit.pxyz, (0, 1, 0, 1, 0, 1))
off = np.rint(unit.pxyz[order]).astype(np.int64)
ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()
cell_nodes = (((ci[:, None] + off[None, :, 0]) * (n + 1) + (cj[:, None] + off[None, :, 1])) * (n + 1) + (ck[:, None] + off[None, :, 2])).astype(np.int32)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    prob = ddm.InterfaceProblem(ctx, cell_nodes, (n + 1) ** 3, S)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record(); v = prob.assemble(S); ev[1].record()
    va = ctx.assemble_cells_bsr(S, prob.cell_nodes, prob.rowptr, prob.colidx); ev[2].record()
    torch.cuda.synchronize()
    print(f"InterfaceProblem (pairs + pattern + plan + assembly) {1e3*(t1-t0):.1f} ms; gather assembly {ev[0].elapsed_time(ev[1]):.2f} ms, "
          f"atomic assembly {ev[1].elapsed_time(ev[2]):.2f} ms, max diff {float((v-va).abs().max()/va.abs().max()):.1e}, nnzb {prob.colidx.numel()}, contributions {prob.plan_contrib.numel()}")
