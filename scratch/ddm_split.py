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
    e0, e1 = ddm.cell_pair_elements(cell_nodes)
    t1 = time.perf_counter()
    d0, d1 = torch.from_numpy(e0).to(ctx.device), torch.from_numpy(e1).to(ctx.device)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    rowptr, colidx = ctx.bsr_pattern(d0, d1, (n + 1) ** 3)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    cn = torch.from_numpy(cell_nodes).to(ctx.device)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    vals = ctx.assemble_cells_bsr(S, cn, rowptr, colidx)
    torch.cuda.synchronize(); t5 = time.perf_counter()
    print(f"pairs(host) {1e3*(t1-t0):.1f} ms ({e0.shape[0]} pairs), upload {1e3*(t2-t1):.1f}, pattern {1e3*(t3-t2):.1f}, upload cell_nodes {1e3*(t4-t3):.1f}, assemble {1e3*(t5-t4):.2f} ms, nnzb {colidx.numel()}")
