"""Dump the BSR pattern of a synthetic lattice for tools/spmv_bench (needs a GPU)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
geom, n, m_ = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lat = M.synthetic_lattice(geom, (n, n, n), [0.05]); mesh = M.mesh_from_synthetic(lat, m_)
ctx = L.Context()
t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)
rp, ci = ctx.bsr_pattern(t(mesh.en0, np.int32), t(mesh.en1, np.int32), mesh.n_nodes)
with open(sys.argv[4], "wb") as f:
    np.array([mesh.n_nodes, ci.numel()], dtype=np.int32).tofile(f); rp.cpu().numpy().tofile(f); ci.cpu().numpy().tofile(f)
print("dumped", mesh.n_nodes, ci.numel())
