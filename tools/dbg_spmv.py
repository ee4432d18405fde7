import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
from pylatticedso_b200.fem import BeamFEM
ctx = L.Context()
lat = M.synthetic_lattice("BCC", (20,20,20), [0.05]); mesh = M.mesh_from_synthetic(lat, 2)
fixed, g, f = M.compression_bc(mesh)
fem = BeamFEM(mesh, 1013.0, 0.3, ctx=ctx); fem.assemble()
fd, gd, fdv = [torch.from_numpy(a).to(ctx.device) for a in (fixed, g, f)]
vbc, b = ctx.apply_dirichlet(fem.rowptr, fem.colidx, fem.vals, fd, gd, fdv)
for dbg in (0, 1, 4, 8):
    for pc in (1, 2):
        u, info = ctx.pcg(fem.rowptr, fem.colidx, vbc, b, tol=1e-8, maxiter=256, precond=pc, profile_iters=128, debug=dbg)
        print(f"debug={dbg} pc={pc} spmv_us={1e3*info['spmv_ms']:.1f} upd_us={1e3*info['update_ms']:.1f}")
# plain spmv warm loop
x = torch.randn(fem.n_dof, dtype=torch.float64, device=ctx.device); y = torch.empty_like(x)
for _ in range(5): ctx.spmv(fem.rowptr, fem.colidx, vbc, x, out=y)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): ctx.spmv(fem.rowptr, fem.colidx, vbc, x, out=y)
e1.record(); torch.cuda.synchronize()
print("plain k_bsr_spmv warm loop us:", 1e3*e0.elapsed_time(e1)/50)
