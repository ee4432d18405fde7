"""One persistent-kernel solve of BASELINE config 1 with the in-kernel phase trace (LAT_PERSIST_TRACE / _FILE)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
from pylatticedso_b200.fem import BeamFEM
ctx = L.Context(); dev = ctx.device
pc = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lat = M.synthetic_lattice("BCC", (20, 20, 20), [0.05]); mesh = M.mesh_from_synthetic(lat, 2)
fixed, g, f = M.compression_bc(mesh)
fem = BeamFEM(mesh, 1013.0, 0.3, ctx=ctx); fem.build_pattern(); fem.assemble()
t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
vbc, b = ctx.apply_dirichlet(fem.rowptr, fem.colidx, fem.vals, t(fixed, np.uint8), t(g, np.float64), t(f, np.float64))
for rep in range(3):
    u, info = ctx.pcg(fem.rowptr, fem.colidx, vbc, b, tol=1e-8, maxiter=200000, precond=pc)
print(f"pc={pc} iters={info['iters']} solve_ms={info['solve_ms']:.3f} us/it={1e3*info['solve_ms']/info['iters']:.2f} persistent={info['persistent']}")
