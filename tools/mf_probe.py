"""configs[1] (BCC 20^3 m=2) solved matrix-free and assembled; prints iteration / kernel times (for ncu captures too)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
from pylatticedso_b200.fem import BeamFEM
geom, n, mseg = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("BCC", 20, 2)
lat = M.synthetic_lattice(geom, (n, n, n), [0.05 if geom == "BCC" else 0.03]); m = M.mesh_from_synthetic(lat, mseg)
fixed, g, f = M.compression_bc(m)
fem = BeamFEM(m, 1013.0, 0.3)
for rep in range(2):
    u, R, i1 = fem.solve_matrix_free(fixed, g, f, tol=1e-8, precond=2, profile_iters=32, want_reactions=False)
    print(f"matfree   iters={i1['iters']} solve={i1['solve_ms']:.2f} ms  {i1['solve_ms']/i1['iters']*1e3:.1f} us/it  spmv={i1['spmv_ms']*1e3:.1f} us update={i1['update_ms']*1e3:.1f} us  {m.n_dof*i1['iters']/i1['solve_ms']/1e6:.2f} G DOF-it/s", flush=True)
if os.environ.get("MF_ONLY"): sys.exit(0)
u, R, i2 = fem.solve(fixed, g, f, tol=1e-8, precond=2, profile_iters=32, want_reactions=False)
print(f"assembled iters={i2['iters']} solve={i2['solve_ms']:.2f} ms  {i2['solve_ms']/i2['iters']*1e3:.1f} us/it  spmv={i2['spmv_ms']*1e3:.1f} us update={i2['update_ms']*1e3:.1f} us  {m.n_dof*i2['iters']/i2['solve_ms']/1e6:.2f} G DOF-it/s", flush=True)
