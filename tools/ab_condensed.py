"""Full solve vs joint-only (strut-condensed) solve at the reference's mesh density (18 elements per strut)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
from pylatticedso_b200.fem import BeamFEM
ctx = L.Context()
for geom, n, mseg, r in (("BCC", 5, 18, 0.05), ("BCC", 10, 18, 0.05), ("BCC", 20, 2, 0.05)):
    lat = M.synthetic_lattice(geom, (n, n, n), [r]); mesh = M.mesh_from_synthetic(lat, mseg)
    fixed, g, f = M.compression_bc(mesh)
    fem = BeamFEM(mesh, 1013.0, 0.3, ctx=ctx)
    out = {}
    for name, fn in (("full mesh, assembled BSR", fem.solve), ("full mesh, matrix-free  ", fem.solve_matrix_free), ("joint-only (condensed)  ", fem.solve_condensed)):
        torch.cuda.synchronize(); fn(fixed, g, f, tol=1e-8)          # warm
        torch.cuda.synchronize(); t0 = time.perf_counter()
        u, R, info = fn(fixed, g, f, tol=1e-8)
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
        out[name] = u[: 6 * mesh.n_points]
        print(f"{geom} {n}^3 m={mseg:2d} {name}: dof={u.numel():8d} iters={info['iters']:6d} pcg={info['solve_ms']:9.2f} ms  whole call (assemble + BC + PCG + reactions, wall)={wall*1e3:9.2f} ms", flush=True)
    ref = out["full mesh, assembled BSR"]
    for k, v in out.items():
        print(f"      {k} vs full assembled at the lattice points: {float((v - ref).abs().max() / ref.abs().max()):.1e}", flush=True)
