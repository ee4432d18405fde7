"""Two-level preconditioner on several GPUs (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/dist_two_level.py [n_big]
1. parity: a small sharded Octet / BCC system with the coarse space (NCCL and peer-memory exchange, assembled and
   matrix-free) against the CPU oracle's direct solve, and the iteration count against the same solve on one GPU;
2. n_big > 0: Octet n_big^3 from the per-slab generator (BASELINE configs[4] at n_big = 100), matrix-free, block-Jacobi
   against two-level: iterations and solve time (max over ranks, CUDA events inside the library)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M, distributed as D  # noqa: E402
from pylatticedso_b200.fem import BeamFEM  # noqa: E402
from oracle import lattice_oracle as orc  # noqa: E402

E, NU = 1013.0, 0.3


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n_big = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = L.Context(local)
    ctx.comm_create(rank, world)
    ok = True
    for geom, n, m_, r, n_agg in (("Octet", (3 * world, 4, 4), 1, 0.03, 3 * world), ("BCC", (2 * world, 3, 3), 2, 0.05, 2 * world)):
        lat = M.synthetic_lattice(geom, n, [r])
        mesh = M.mesh_from_synthetic(lat, m_)
        fixed, g, f = M.compression_bc(mesh)
        f = f.copy(); f[6 * 7 + 0] = 0.01
        dfem = D.DistributedFEM(ctx, mesh, E, NU, rank, world)
        dfem.set_bc(fixed, g, f)
        K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E, NU)
        uo, Ro = orc.solve_static(K, fixed.astype(bool), g, f)
        # the same system with the same number of boxes on ONE GPU (rank 0): the coarse space is the same space,
        # so the iteration counts must agree to rounding
        it1 = None
        if rank == 0:
            c1 = L.Context(local)
            fem1 = BeamFEM(mesh, E, NU, ctx=c1)
            _, _, i1 = fem1.solve(fixed, g, f, tol=1e-12, two_level=n_agg)
            _, _, i0 = fem1.solve(fixed, g, f, tol=1e-12, persistent=False)
            it1 = (i1["iters"], i0["iters"])
            del fem1
            c1.close()
        for mode in ("nccl", "p2p"):
            if mode == "p2p":
                dfem.enable_p2p()
            for op in ("assembled", "matfree"):
                solve = dfem.solve if op == "assembled" else dfem.solve_matrix_free
                for rep in range(2):
                    u, R, info = solve(tol=1e-12, maxiter=100000, precond=L.PC_BLOCK6, check_every=16, two_level=n_agg)
                ug, Rg = dfem.gather_owned(u), dfem.gather_owned(R)
                if rank == 0:
                    eu = np.abs(ug - uo).max() / np.abs(uo).max()
                    er = np.abs(Rg - Ro).max() / np.abs(Ro).max()
                    print(f"[dist_two_level] {mode} {op} {geom}{n} m={m_} world={world} n_dof={mesh.n_dof} n_agg={n_agg} iters={info['iters']} "
                          f"(one GPU: two-level {it1[0]}, block-Jacobi {it1[1]}) info={info['info']} two_level={info['two_level']} "
                          f"graph={info['graph']} true_relres={info['true_relres']:.1e} |u-uo|/|uo|={eu:.2e} |R-Ro|/|Ro|={er:.2e}", flush=True)
                    ok = ok and info["info"] == 0 and info["two_level"] and eu < 1e-8 and er < 1e-8 and abs(info["iters"] - it1[0]) <= 3
        ctx.p2p_destroy()
        del dfem
        # the joint-only (strut-condensed) system, sharded, with the coarse space over the joints
        if m_ > 1:
            jf = D.DistributedJointFEM(ctx, mesh, E, NU, rank, world)
            nj = 6 * mesh.n_points
            jf.set_bc(fixed[:nj], g[:nj], f[:nj])
            jf.enable_p2p()
            uj, Rj, info = jf.solve(tol=1e-12, maxiter=100000, precond=L.PC_BLOCK6, two_level=n_agg)
            uj0, _, info0 = jf.solve(tol=1e-12, maxiter=100000, precond=L.PC_BLOCK6, persistent=False)
            ujg = jf.gather_owned(uj)
            if rank == 0:
                eu = np.abs(ujg - uo[:nj]).max() / np.abs(uo).max()
                print(f"[dist_two_level] joint-only {geom}{n} m={m_} world={world} n_dof={nj} iters={info['iters']} (block-Jacobi {info0['iters']}) "
                      f"info={info['info']} two_level={info['two_level']} |u-uo|/|uo|={eu:.2e}", flush=True)
                ok = ok and info["info"] == 0 and info["two_level"] and eu < 1e-8
            ctx.p2p_destroy()
            del jf
    if n_big > 0:
        t0 = time.perf_counter()
        dfem = D.DistributedFEM.from_generator(ctx, "Octet", (n_big,) * 3, [0.03], 1, E, NU, rank, world)
        fx, gg, ff = D.compression_bc_local(dfem.lmesh)
        dfem.set_bc_local(fx, gg, ff)
        dfem.enable_p2p()
        torch.cuda.synchronize(); dist.barrier()
        t_gen = time.perf_counter() - t0
        u0, _, i0 = dfem.solve_matrix_free(tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, want_reactions=False)
        for rep in range(2):
            torch.cuda.synchronize(); dist.barrier()
            t1 = time.perf_counter()
            tl = dfem.two_level()
            torch.cuda.synchronize(); dist.barrier()
            t_setup = time.perf_counter() - t1
        u2, _, i2 = dfem.solve_matrix_free(tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, want_reactions=False, two_level=tl)
        no = 6 * dfem.n_owned
        dd = torch.stack([(u2[:no] - u0[:no]).abs().max(), u0[:no].abs().max()])
        tt = torch.tensor([i0["solve_ms"], i2["solve_ms"], 1e3 * t_setup], dtype=torch.float64, device=ctx.device)
        dist.all_reduce(dd, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"[dist_two_level] Octet {n_big}^3 ({dfem.n_dof_global} DOF) world={world} matrix-free: block-Jacobi {i0['iters']} it "
                  f"{float(tt[0]):.1f} ms | two-level ({tl.n_agg} aggregates) {i2['iters']} it {float(tt[1]):.1f} ms + set-up {float(tt[2]):.1f} ms "
                  f"| info {i0['info']}/{i2['info']} true_relres {i2['true_relres']:.1e} |du|/|u| {float(dd[0] / dd[1]):.1e} "
                  f"(generate + upload {t_gen:.2f} s)", flush=True)
            ok = ok and i2["info"] == 0 and float(dd[0] / dd[1]) < 1e-5
        ctx.p2p_destroy()
    ctx.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("[dist_two_level] PASS" if ok else "[dist_two_level] FAIL", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
