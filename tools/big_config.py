"""BASELINE configs 3-5 at full size: assemble + PCG to 1e-8 (and the adjoint gradient for config 3).
   single GPU:  python tools/big_config.py octet100
   N GPUs:      python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/big_config.py octet100
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M

E, NU = 1013.0, 0.3
CFG = {"octet100": ("Octet", 100, 0.03, None), "octet40": ("Octet", 40, 0.03, ("linear", [False, False, True], [0, 0, 0.0125])),
       "bcc60": ("BCC", 60, 0.05, None), "octet74": ("Octet", 74, 0.03, None)}
name = sys.argv[1] if len(sys.argv) > 1 else "octet100"
geom, n, r, grad = CFG[name]
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = L.Context(local)
t0 = time.time()
lat = M.synthetic_lattice(geom, (n, n, n), [r], grad_radius=grad)
mesh = M.mesh_from_synthetic(lat, 1)
fixed, g, f = M.compression_bc(mesh)
t_host = time.time() - t0
peak = 6554.6
ev = lambda: torch.cuda.Event(enable_timing=True)
out = {"config": name, "geom": geom, "n": n, "n_gpus": world, "n_nodes": mesh.n_nodes, "n_elements": mesh.n_elems, "n_dof": mesh.n_dof,
       "host_mesh_generation_s": round(t_host, 2)}
if world == 1:
    from pylatticedso_b200.fem import BeamFEM
    fem = BeamFEM(mesh, E, NU, ctx=ctx)
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record(); fem.build_pattern(); b.record(); fem.assemble(); c.record(); torch.cuda.synchronize()
    out.update(pattern_ms=a.elapsed_time(b), assemble_ms=b.elapsed_time(c), nnzb=fem.nnzb,
               assembly_elements_per_s=mesh.n_elems / (b.elapsed_time(c) * 1e-3), matrix_GB=fem.nnzb * 288 / 1e9)
    dev = ctx.device
    fd, gd, fv = [torch.from_numpy(v).to(dev) for v in (fixed, g, f)]
    vbc, rhs = ctx.apply_dirichlet(fem.rowptr, fem.colidx, fem.vals, fd, gd, fv)
    u, info = ctx.pcg(fem.rowptr, fem.colidx, vbc, rhs, tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, profile_iters=32)
    nn, nz = fem.n_nodes, fem.nnzb
    out.update(pcg=info, dof_iters_per_s=mesh.n_dof * info["iters"] / (info["solve_ms"] * 1e-3),
               spmv_GBps=(nz * 292 + nn * 148) / info["spmv_ms"] / 1e6, spmv_frac=(nz * 292 + nn * 148) / info["spmv_ms"] / 1e6 / peak,
               iteration_frac=(nz * 292 + nn * 100 + 6 * nn * 144) * info["iters"] / info["solve_ms"] / 1e6 / peak)
    # the same system without an assembled matrix
    fem.vals = None; del vbc; torch.cuda.empty_cache()
    um, Rm, im = fem.solve_matrix_free(fixed, g, f, tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, profile_iters=32, want_reactions=False)
    ctx.set_dirichlet_values(fd, gd, u)
    out.update(matrix_free=dict(pcg=im, dof_iters_per_s=mesh.n_dof * im["iters"] / (im["solve_ms"] * 1e-3),
                                rel_diff_vs_assembled=float((um - u).abs().max() / u.abs().max())))
    fem.assemble()
    if name == "octet40":   # config 3: adjoint compliance gradient w.r.t. the 64 000 cell radii
        ctx.set_dirichlet_values(fd, gd, u)
        grp = torch.from_numpy(mesh.cell_of_elem.astype(np.int32)).to(dev)
        e0, e1 = ev(), ev(); e0.record()
        gr = ctx.compliance_grad(fem.x, fem.y, fem.z, fem.en0, fem.en1, fem.rad, grp, n ** 3, u, E, NU)
        e1.record(); torch.cuda.synchronize()
        out.update(gradient_ms=e0.elapsed_time(e1), gradient_elements_per_s=mesh.n_elems / (e0.elapsed_time(e1) * 1e-3),
                   gradient_norm=float(gr.norm()))
else:
    from pylatticedso_b200 import distributed as D
    ctx.comm_create(rank, world)
    dfem = D.DistributedFEM(ctx, mesh, E, NU, rank, world)
    dfem.set_bc(fixed, g, f)
    mode = "nccl"
    if "--nccl" not in sys.argv:
        dfem.enable_p2p(); mode = "nvlink-peer-memory"
    a, b = ev(), ev(); a.record(); dfem.assemble(); b.record(); torch.cuda.synchronize()
    u, R, info = dfem.solve(tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6)
    u, R, info = dfem.solve(tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6)     # second solve: warm
    t = torch.tensor([info["solve_ms"], b.elapsed_time(a) * -1.0], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nz = torch.tensor([float(dfem.nnzb_owned), float(dfem.n_owned)], dtype=torch.float64, device=ctx.device); dist.all_reduce(nz)
    nzb, nn = float(nz[0]), float(nz[1])
    out.update(exchange=mode, pcg=info, solve_ms_max=float(t[0]), assemble_ms_max=float(t[1]),
               dof_iters_per_s=mesh.n_dof * info["iters"] / (float(t[0]) * 1e-3),
               iteration_frac_of_aggregate_hbm=(nzb * 292 + nn * 100 + 6 * nn * 144) * info["iters"] / float(t[0]) / 1e6 / (peak * world),
               ghosts_per_rank=int(dfem.n_local - dfem.n_owned))
    um, Rm, im = dfem.solve_matrix_free(tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, want_reactions=False)
    um, Rm, im = dfem.solve_matrix_free(tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, want_reactions=False)
    tm = torch.tensor([im["solve_ms"]], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dmax = ((um - u).abs().max() / u.abs().max()).reshape(1); dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
    out.update(matrix_free=dict(pcg=im, solve_ms_max=float(tm[0]), dof_iters_per_s=mesh.n_dof * im["iters"] / (float(tm[0]) * 1e-3),
                                rel_diff_vs_assembled=float(dmax[0])))
    ctx.p2p_destroy(); ctx.comm_destroy()
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
