"""Multi-GPU parity check (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py
The slab-sharded assemble + PCG (NCCL halo exchange, all-reduced dots) must reproduce the CPU
oracle's displacements and reactions to 1e-8."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M, distributed as D  # noqa: E402
from oracle import lattice_oracle as orc  # noqa: E402

E, NU = 1013.0, 0.3


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = L.Context(local)
    ctx.comm_create(rank, world)
    ok = True
    for geom, n, m_, r in (("BCC", (2 * world, 3, 3), 2, 0.05), ("Octet", (2 * world + 1, 2, 3), 1, 0.03)):
        lat = M.synthetic_lattice(geom, n, [r])
        mesh = M.mesh_from_synthetic(lat, m_)
        fixed, g, f = M.compression_bc(mesh)
        f = f.copy(); f[6 * 7 + 0] = 0.01
        dfem = D.DistributedFEM(ctx, mesh, E, NU, rank, world)
        dfem.set_bc(fixed, g, f)
        K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E, NU)     # every rank: the joint-only
        uo, Ro = orc.solve_static(K, fixed.astype(bool), g, f)                                  # check compares locally
        # p2p-fused: the persistent on-chip kernel (halo + all-reduce inside its grid barriers) where the slabs fit the
        # shared memory, i.e. here; p2p-fused-3k: the same exchange in the three-kernel iteration
        for mode in ("nccl", "p2p", "p2p-fused", "p2p-fused-3k"):
            if mode == "p2p":
                dfem.enable_p2p()
            kw = dict(fused_halo=mode.startswith("p2p-fused"))
            for op in ("assembled", "matfree"):
                if op == "matfree" and mode == "p2p-fused-3k":
                    continue
                solve = dfem.solve if op == "assembled" else dfem.solve_matrix_free
                kw.pop("persistent", None)
                if op == "assembled":
                    kw["persistent"] = mode != "p2p-fused-3k"
                for rep in range(1 if mode == "nccl" else 2):      # second p2p solve: flags of the first must not match
                    u, R, info = solve(tol=1e-13, maxiter=100000, precond=L.PC_BLOCK6, check_every=16, **kw)
                ug = dfem.gather_owned(u)
                Rg = dfem.gather_owned(R)
                if rank == 0:
                    eu = np.abs(ug - uo).max() / np.abs(uo).max()
                    er = np.abs(Rg - Ro).max() / np.abs(Ro).max()
                    print(f"[dist_check] {mode} {op} {geom}{n} m={m_} world={world} n_dof={mesh.n_dof} iters={info['iters']} "
                          f"info={info['info']} persistent={info.get('persistent', False)} relres={info['relres']:.1e} solve_ms={info['solve_ms']:.2f} |u-uo|/|uo|={eu:.2e} "
                          f"|R-Ro|/|Ro|={er:.2e}", flush=True)
                    ok = ok and info["info"] in (0, 5) and eu < 1e-8 and er < 1e-8
                    if mode == "p2p-fused" and op == "assembled":
                        ok = ok and bool(info.get("persistent"))
        # sharded compliance gradient w.r.t. per-cell radii == oracle's
        ncell = int(mesh.cell_of_elem.max()) + 1
        gd = dfem.compliance_gradient(u, mesh.cell_of_elem, ncell).cpu().numpy()
        if rank == 0:
            go = orc.compliance_gradient(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, uo, mesh.cell_of_elem, ncell, E, NU)
            eg = np.abs(gd - go).max() / np.abs(go).max()
            print(f"[dist_check] gradient {geom}{n} world={world} n_groups={ncell} |g-go|/|go|={eg:.2e}", flush=True)
            ok = ok and eg < 1e-6
        ctx.p2p_destroy()
        # joint-only (strut-condensed) system, sharded: joints == the oracle's FULL solve at the lattice points, and the
        # back-substituted interior nodes of every rank's struts == the oracle's full field
        if m_ > 1:
            jf = D.DistributedJointFEM(ctx, mesh, E, NU, rank, world)
            nj = 6 * mesh.n_points
            jf.set_bc(fixed[:nj], g[:nj], f[:nj])
            jf.enable_p2p()
            for persistent in (True, False):
                uj, Rj, info = jf.solve(tol=1e-13, maxiter=100000, precond=L.PC_BLOCK6, persistent=persistent)
                ujg, Rjg = jf.gather_owned(uj), jf.gather_owned(Rj)
                uf = jf.recover_full_field(uj).cpu().numpy().reshape(-1, 6)
                nodes = np.concatenate([jf.part.local_nodes, jf.full["interior_global"]])
                e_int = np.abs(uf - uo.reshape(-1, 6)[nodes]).max() / np.abs(uo).max()
                e_all = torch.tensor([e_int], dtype=torch.float64, device=ctx.device)
                dist.all_reduce(e_all, op=dist.ReduceOp.MAX)
                if rank == 0:
                    eu = np.abs(ujg - uo[:nj]).max() / np.abs(uo).max()
                    er = np.abs(Rjg - Ro[:nj]).max() / np.abs(Ro).max()
                    print(f"[dist_check] joint-only {geom}{n} m={m_} world={world} n_dof={nj} (full {mesh.n_dof}) iters={info['iters']} "
                          f"info={info['info']} persistent={info.get('persistent', False)} solve_ms={info['solve_ms']:.2f} "
                          f"|u-uo|/|uo|={eu:.2e} |R-Ro|/|Ro|={er:.2e} full field (max over ranks) {float(e_all):.2e}", flush=True)
                    ok = ok and info["info"] in (0, 5) and eu < 1e-8 and er < 1e-8 and float(e_all) < 1e-8
            ctx.p2p_destroy()
    ctx.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("[dist_check] PASS" if ok else "[dist_check] FAIL", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
