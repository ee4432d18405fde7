"""Where the end-to-end time of BeamFEM(mesh) + BeamFEM.solve goes (config 1), stage by stage (host clock, synced)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
from pylatticedso_b200.fem import BeamFEM
ctx = L.Context(); dev = ctx.device
lat = M.synthetic_lattice("BCC", (20, 20, 20), [0.05]); mesh = M.mesh_from_synthetic(lat, 2)
fixed, g, f = M.compression_bc(mesh)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
keep = {k: pin(getattr(mesh, k)) for k in ("x", "y", "z", "en0", "en1", "rad")}
pm = M.BeamMesh(**{k: v.numpy() for k, v in keep.items()}, beam_of_elem=mesh.beam_of_elem, chain=mesh.chain, n_points=mesh.n_points,
                point_index=mesh.point_index, cell_of_elem=mesh.cell_of_elem)
bc = [pin(v) for v in (fixed, g, f)]
uh = torch.empty(mesh.n_dof, dtype=torch.float64).pin_memory(); Rh = torch.empty_like(uh).pin_memory()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    t0 = T(); fem = BeamFEM(pm, 1013.0, 0.3, ctx=ctx); t1 = T()
    fem.build_pattern(); t2 = T()
    fem.assemble(); t3 = T()
    fd, gd, fv = [torch.as_tensor(b.numpy()).to(dev) for b in bc]; t4 = T()
    vbc, b = ctx.apply_dirichlet(fem.rowptr, fem.colidx, fem.vals, fd, gd, fv); t5 = T()
    u, info = ctx.pcg(fem.rowptr, fem.colidx, vbc, b, tol=1e-8, maxiter=200000, precond=2); t6 = T()
    ctx.set_dirichlet_values(fd, gd, u); R = ctx.spmv(fem.rowptr, fem.colidx, fem.vals, u); t7 = T()
    uh.copy_(u, non_blocking=True); Rh.copy_(R, non_blocking=True); t8 = T()
    print(f"rep {rep}: upload {1e3*(t1-t0):.2f} | pattern {1e3*(t2-t1):.2f} | assemble {1e3*(t3-t2):.2f} | bc upload {1e3*(t4-t3):.2f} | dirichlet {1e3*(t5-t4):.2f} | "
          f"pcg call {1e3*(t6-t5):.2f} (solve_ms {info['solve_ms']:.2f}, persistent={info['persistent']}) | reactions {1e3*(t7-t6):.2f} | D2H {1e3*(t8-t7):.2f} | total {1e3*(t8-t0):.2f} ms", flush=True)
    del fem, vbc, b, u, R
