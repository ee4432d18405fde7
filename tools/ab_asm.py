"""A/B of the three fused assembly mappings (CUDA events, 3 warm-up + 10 timed, matrices larger than L2)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
E, NU = 1013.0, 0.3
ctx = L.Context(); dev = ctx.device
t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for geom, n, m_, r in (("BCC", 60, 1, 0.05), ("BCC", 40, 2, 0.05), ("Octet", 40, 1, 0.03)):
    lat = M.synthetic_lattice(geom, (n, n, n), [r]); mesh = M.mesh_from_synthetic(lat, m_)
    x, y, z, en0, en1, rad = t(mesh.x, np.float64), t(mesh.y, np.float64), t(mesh.z, np.float64), t(mesh.en0, np.int32), t(mesh.en1, np.int32), t(mesh.rad, np.float64)
    Ec, N = mesh.n_elems, mesh.n_nodes
    rowptr, colidx = ctx.bsr_pattern(en0, en1, N); nnzb = colidx.numel()
    vals = torch.empty(nnzb * 36, dtype=torch.float64, device=dev)
    ref = None
    for name, mode in (("gather", L.ASM_GATHER), ("rows", L.ASM_ROWS), ("atomic", L.ASM_ATOMIC)):
        ms = timeit(lambda: ctx.assemble_bsr(x, y, z, en0, en1, rad, N, nnzb, E, NU, mode=mode, out=vals))
        if ref is None: ref = vals.clone(); err = 0.0
        else: err = float((vals - ref).abs().max() / ref.abs().max())
        print(f"{geom} {n}^3 m={m_} E={Ec} nnzb={nnzb} {name:7s} {ms*1e3:9.1f} us  {Ec/ms/1e6:7.2f} G elem/s  write {nnzb*288/ms/1e6:6.0f} GB/s  maxrel vs gather {err:.1e}", flush=True)
