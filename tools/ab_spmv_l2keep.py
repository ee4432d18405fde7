"""Three-kernel PCG iteration with / without the L2-resident matrix part (k_cg_spmv<.., .., true>, LAT_SPMV_L2KEEP)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
from pylatticedso_b200.fem import BeamFEM
ctx = L.Context(); dev = ctx.device
t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
for geom, n, m_, r in (("BCC", (20, 20, 20), 2, 0.05), ("BCC", (24, 24, 24), 2, 0.05), ("BCC", (32, 32, 32), 2, 0.05), ("Octet", (40, 40, 40), 1, 0.03)):
    lat = M.synthetic_lattice(geom, n, [r]); mesh = M.mesh_from_synthetic(lat, m_)
    fixed, g, f = M.compression_bc(mesh)
    fem = BeamFEM(mesh, 1013.0, 0.3, ctx=ctx); fem.build_pattern(); fem.assemble()
    vbc, b = ctx.apply_dirichlet(fem.rowptr, fem.colidx, fem.vals, t(fixed, np.uint8), t(g, np.float64), t(f, np.float64))
    mb = fem.nnzb * 288 / 1e6
    res = {}
    for keep in ("0", None):
        if keep is None: os.environ.pop("LAT_SPMV_L2KEEP", None)
        else: os.environ["LAT_SPMV_L2KEEP"] = keep
        for rep in range(2):
            u, info = ctx.pcg(fem.rowptr, fem.colidx, vbc, b, tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6, persistent=False)
        res[keep] = (info, u.clone())
        print(f"{geom}{n} m={m_} ndof={mesh.n_dof} matrix {mb:7.1f} MB  L2KEEP={'default' if keep is None else keep:7s} iters={info['iters']} "
              f"solve_ms={info['solve_ms']:.3f} us/it={1e3*info['solve_ms']/info['iters']:.2f}", flush=True)
    d = float((res["0"][1] - res[None][1]).abs().max() / res["0"][1].abs().max())
    print(f"   same iterates: iters {res['0'][0]['iters']} / {res[None][0]['iters']}, max rel diff {d:.1e}")
    del fem, vbc, b
    torch.cuda.empty_cache()
