import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
from pylatticedso_b200.fem import BeamFEM
E, nu = 1013.0, 0.3
ctx = L.Context()
for geom, n, m_ in (("BCC", (20,20,20), 2), ("BCC",(60,60,60),1), ("Octet",(40,40,40),1)):
    lat = M.synthetic_lattice(geom, n, [0.05 if geom=="BCC" else 0.03]); mesh = M.mesh_from_synthetic(lat, m_)
    fixed, g, f = M.compression_bc(mesh)
    fem = BeamFEM(mesh, E, nu, ctx=ctx); fem.assemble()
    dev = ctx.device
    fd, gd, fdv = [torch.from_numpy(a).to(dev) for a in (fixed, g, f)]
    vbc, b = ctx.apply_dirichlet(fem.rowptr, fem.colidx, fem.vals, fd, gd, fdv)
    for variant in ("classic", "cg"):
        for pc in (1, 2):
            u, info = ctx.pcg(fem.rowptr, fem.colidx, vbc, b, tol=1e-8, maxiter=5000, precond=pc, classic=(variant == "classic"), profile_iters=0)
            byt = fem.nnzb*292 + fem.n_nodes*(4+(4 if variant == "classic" else 3)*48)
            r = fem.ctx.spmv(fem.rowptr, fem.colidx, vbc, u) - b
            print(f"{geom}{n} m={m_} ndof={mesh.n_dof} nnzb={fem.nnzb} {variant} pc={pc} iters={info['iters']} info={info['info']} relres={info['relres']:.2e} true={float(r.norm())/info['norm_b']:.2e} "
                  f"solve_ms={info['solve_ms']:.2f} us/iter={1e3*info['solve_ms']/max(1,info['iters']):.1f} restarts={info['restarts']} true={info['true_relres']:.2e}")
    del fem, vbc, b
