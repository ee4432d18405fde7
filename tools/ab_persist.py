"""A/B: persistent on-chip PCG kernel (csrc/pcg_persist.cuh) vs the three-kernel iteration on BASELINE config 1
and a few other sizes.  Prints iterations, solve time, us/iteration, G DOF-it/s and the difference of the iterates."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
from pylatticedso_b200.fem import BeamFEM
E, NU = 1013.0, 0.3
ctx = L.Context(); dev = ctx.device
cases = [("BCC", (20, 20, 20), 2, 0.05), ("BCC", (12, 12, 12), 2, 0.05), ("Octet", (12, 12, 12), 1, 0.03), ("BCC", (24, 24, 24), 2, 0.05)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for geom, n, m_, r in cases:
    lat = M.synthetic_lattice(geom, n, [r]); mesh = M.mesh_from_synthetic(lat, m_)
    fixed, g, f = M.compression_bc(mesh)
    fem = BeamFEM(mesh, E, NU, ctx=ctx)
    fem.build_pattern(); fem.assemble()
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
    fd, gd, fv = t(fixed, np.uint8), t(g, np.float64), t(f, np.float64)
    vbc, b = ctx.apply_dirichlet(fem.rowptr, fem.colidx, fem.vals, fd, gd, fv)
    for pc in (L.PC_BLOCK6, L.PC_JACOBI):
        out = {}
        for name, kw in (("three-kernel", dict(persistent=False)), ("persistent", dict(persistent=True))):
            best = None
            for rep in range(4):
                u, info = ctx.pcg(fem.rowptr, fem.colidx, vbc, b, tol=1e-8, maxiter=200000, precond=pc, **kw)
                if best is None or info["solve_ms"] < best["solve_ms"]:
                    best = info
            out[name] = (u.clone(), best)
            print(f"{geom}{n} m={m_} ndof={mesh.n_dof} pc={pc} {name:13s} persistent={best.get('persistent')} iters={best['iters']} info={best['info']} "
                  f"solve_ms={best['solve_ms']:.3f} us/it={1e3 * best['solve_ms'] / max(best['iters'], 1):.2f} "
                  f"GDOFit/s={mesh.n_dof * best['iters'] / best['solve_ms'] / 1e6:.2f} true_relres={best['true_relres']:.2e} restarts={best['restarts']}", flush=True)
        d = float((out["persistent"][0] - out["three-kernel"][0]).abs().max() / out["three-kernel"][0].abs().max())
        print(f"   max |u_persistent - u_three_kernel| / max|u| = {d:.2e}", flush=True)
    # tight tolerance: both against each other at 1e-12
    ua, ia = ctx.pcg(fem.rowptr, fem.colidx, vbc, b, tol=1e-12, maxiter=400000, precond=L.PC_BLOCK6, persistent=False)
    ub, ib = ctx.pcg(fem.rowptr, fem.colidx, vbc, b, tol=1e-12, maxiter=400000, precond=L.PC_BLOCK6, persistent=True)
    print(f"   tol 1e-12: iters {ia['iters']} / {ib['iters']}  info {ia['info']} / {ib['info']}  diff {float((ua - ub).abs().max() / ua.abs().max()):.2e}  "
          f"repro: {bool(torch.equal(ub, ctx.pcg(fem.rowptr, fem.colidx, vbc, b, tol=1e-12, maxiter=400000, precond=L.PC_BLOCK6, persistent=True)[0]))}", flush=True)
