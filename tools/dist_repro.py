"""Run the slab-sharded bench problem several times; iteration counts must be identical (determinism)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M, distributed as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = L.Context(local); ctx.comm_create(rank, world)
lat = M.synthetic_lattice("BCC", (20 * world, 20, 20), [0.05]); mesh = M.mesh_from_synthetic(lat, 2)
fixed, g, f = M.compression_bc(mesh)
dfem = D.DistributedFEM(ctx, mesh, 1013.0, 0.3, rank, world); dfem.set_bc(fixed, g, f)
for mode, kw in (("nccl", {}), ("p2p-sequential", dict(overlap=False, fused_halo=False)), ("p2p-overlap", dict(overlap=True, fused_halo=False)),
                 ("p2p-fused", dict(fused_halo=True)), ("p2p-fused-3k", dict(fused_halo=True, persistent=False))):
    if mode == "p2p-sequential": dfem.enable_p2p()
    for op, solve in (("assembled", dfem.solve), ("matfree", dfem.solve_matrix_free)):
        if op == "matfree" and "persistent" in kw:
            continue
        its = []
        for rep in range(4):
            u, R, info = solve(tol=1e-8, maxiter=200000, precond=2, **kw)
            its.append((info["iters"], round(info["solve_ms"], 2), round(1e3 * info["solve_ms"] / info["iters"], 1),
                        float(u[: 6 * dfem.n_owned].double().abs().sum())))
        same = len({(i[0], i[3]) for i in its}) == 1
        if rank == 0: print(mode, op, "persistent" if info.get("persistent") else "", "bit-identical runs" if same else "RUNS DIFFER",
                            "(iters, solve ms, us/iter, checksum)", its, flush=True)
ctx.p2p_destroy(); ctx.comm_destroy(); dist.barrier(); dist.destroy_process_group()
