#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 beam-FEM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): BCC 20x20x20 lattice, r = 0.05, 2 elements per
strut (81 261 nodes / 128 000 elements / 487 566 DOF), uniaxial compression
(clamp Zmin, u_z = -0.01 on Zmax), VeroClear.  One STEP = one pass of the hot
path: fused element generation + BSR assembly -> Dirichlet elimination ->
6x6 block-Jacobi PCG to 1e-8 relative residual -> reactions.

value  = PCG DOF-iterations / s = n_dof * iterations / device time of the WHOLE step
         (assembly, elimination and reactions included), inputs resident in HBM.
e2e    = the same quantity through the host-facing API (host numpy buffers in pinned
         memory -> H2D of the mesh and BCs, pattern reuse, step, D2H of u and R).
roofline = the dominant kernel (k_cg_spmv: BSR SpMV fused with the three CG dot products),
         timed live with CUDA events on the library stream during the timed steps.

Multi-GPU (N > 1): weak scaling, one process per GPU; the lattice is extended to
(20 N) x 20 x 20 cells and cut into N x-slabs (see pylatticedso_b200/distributed.py).

Every line also carries
  parity   N > 1: the sharded solve (assembled AND matrix-free, peer-memory exchange) at tol 1e-12 against the
           same system solved on rank 0 alone (u, reactions, compliance gradient); N = 1: assembled vs
           matrix-free at full size + the CUDA path against the CPU oracle's direct solve on a BCC 6^3 case.
           The process exits non-zero when a figure is above 1e-8 (u, R) / 1e-6 (gradient).
  config5  BASELINE configs[4]: Octet 100^3 (24.4 M DOF), r = 0.03, strong scaling over the N GPUs, generated
           per slab on each rank; assembled and matrix-free solve to 1e-8, time to first iteration, host RSS;
           matrix_free_two_level: the same matrix-free solve with the two-level preconditioner (csrc/coarse.cuh).
  ddm_config3 (N = 1)  BASELINE configs[3] through the DDM path, block-Jacobi and two-level interface solve.
The reference arm (--impl reference) runs the SAME (20 N) x 20 x 20 workload with Jacobi-PCG (C/OpenMP oracle
port) on all host threads of the box (the thread count is set explicitly: torchrun exports OMP_NUM_THREADS=1).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

E_MOD, NU, KAPPA = 1013.0, 0.3, 0.9
METRIC = "pcg_dof_iters_per_s"
UNIT = "DOF-iterations/s"
# ncu --set full capture of k_cg_spmv on this workload (profiles/r01_ncu_full_v4_kernels.txt):
# dram__bytes_read.sum 106.76 MB + dram__bytes_write.sum 3.65 MB per launch (algorithmic: 110.5 MB)
NCU_TRAFFIC_CG_SPMV = 111.3e6   # dram__bytes_read.sum + dram__bytes_write.sum of one k_cg_spmv launch (profiles/r01_ncu_full_v4_kernels.txt)
NCU_TRAFFIC_PERSIST_PER_ITER = 40.3e6    # (dram__bytes_read.sum + dram__bytes_write.sum) / 692 iterations of one k_pcg_persist launch (profiles/r02_ncu_persist.txt)
WORKLOAD = "BCC 20x20x20, r=0.05, 2 elements/strut (487566 DOF), uniaxial compression, assemble + PCG to 1e-8"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            c = [v.strip() for v in row.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(n_slabs=1):
    from pylatticedso_b200 import mesh as M
    lat = M.synthetic_lattice("BCC", (20 * n_slabs, 20, 20), [0.05])
    mesh = M.mesh_from_synthetic(lat, 2)
    fixed, g, f = M.compression_bc(mesh)
    return lat, mesh, fixed, g, f


def spmv_bytes(n_nodes, nnzb):
    """Algorithmic bytes of one k_cg_spmv launch: matrix 288 B + column index 4 B per
    block; per block row 4 B rowptr, 48 B u (the gathered vector, counted once) + 48 B r read,
    48 B w = A u written."""
    return nnzb * 292 + n_nodes * (4 + 3 * 48)


def iteration_bytes(n_nodes, nnzb, block_jacobi=True):
    """SpMV (292 nb + 100)/6 B/DOF + 96 B/DOF, + 28 B/DOF for the 6x6 block-Jacobi inverse
    (stored as its 21 unique entries per node; SURVEY budgeted 48 B/DOF for a full 36-entry block)."""
    return nnzb * 292 + n_nodes * 100 + 6 * n_nodes * (96 + (28 if block_jacobi else 0))


# --------------------------------------------------------------------------- reference arm / CPU baseline
def cpu_reference_run(steps, warmup, pcg_iters=3000, n_slabs=1):
    """The reference's CPU path for this workload, restated (oracle port, C + OpenMP on all host threads,
    oracle/oracle_c.c): element matrices, value assembly into the CSR pattern (pattern from scipy COO->CSR of
    the element connectivity), Dirichlet elimination, then the reference's PCG
    (conjugate_gradient_solver.py semantics, JACOBI M = diag(K)^-1) for `pcg_iters` iterations per step (bounded
    sample).  The reference's own solver file cannot travel to the GPU box (no checkout there), hence the port."""
    from oracle import lattice_oracle as orc
    from oracle import oracle_c as oc
    import scipy.sparse as sp
    _, mesh, fixed, g, f = build_workload(n_slabs)
    en = np.stack([mesh.en0, mesh.en1], 1)
    xyz = mesh.xyz
    threads = oc.set_num_threads(oc.host_threads())     # not the inherited OMP_NUM_THREADS (torchrun sets it to 1)
    # one-off pattern (not timed, like the GPU arm's pattern build)
    dofs = (en[:, :, None] * 6 + np.arange(6)[None, None, :]).reshape(-1, 12)
    P = sp.coo_matrix((np.ones(dofs.shape[0] * 144, dtype=np.int8), (np.repeat(dofs, 12, 1).ravel(), np.tile(dofs, (1, 12)).ravel())),
                      shape=(mesh.n_dof, mesh.n_dof)).tocsr()
    P.sum_duplicates(); P.sort_indices()
    indptr, indices = P.indptr.astype(np.int32), P.indices.astype(np.int32)
    del P, dofs
    times, asm_times = [], []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        Ke = oc.elem_stiffness(xyz, en, mesh.rad, E_MOD, NU)
        data = oc.assemble_csr_values(en, Ke, indptr, indices)
        t1 = time.perf_counter()
        K = sp.csr_matrix((data, indices, indptr), shape=(mesh.n_dof, mesh.n_dof))
        Kbc, b = orc.apply_dirichlet(K, fixed, g, f)
        Kbc.sort_indices()
        x, info, it = oc.pcg(Kbc.indptr, Kbc.indices, Kbc.data, b, 1.0 / Kbc.diagonal(), maxiter=pcg_iters, tol=1e-8,
                             mintol=0.0, restart_every=0, alpha_max=1e300)
        t2 = time.perf_counter()
        if s >= warmup:
            times.append(t2 - t0)
            asm_times.append(t1 - t0)
    T = float(np.sum(times))
    val = mesh.n_dof * it * steps / T
    return dict(value=val, ms_per_step=1e3 * T / steps, asm_elems_per_s=mesh.n_elems * steps / float(np.sum(asm_times)),
                n_dof=mesh.n_dof, n_elems=mesh.n_elems, iters=it, threads=threads, converged=(info == 0))


def workload_name(world):
    return WORKLOAD if world == 1 else (f"BCC {20 * world}x20x20 in {world} x-slabs of 20 cell layers "
                                        f"(weak scaling of: {WORKLOAD})")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    r = cpu_reference_run(steps, warmup, n_slabs=args.gpus)
    sample = (f"{steps} step(s): C/OpenMP assembly of the full {r['n_elems']}-element mesh + Jacobi-PCG to 1e-8 capped at 3000 "
              f"iterations ({r['iters']} run) per step, {r['threads']} threads")
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "n_dof": r["n_dof"],
                   "n_elements": r["n_elems"], "precond": "jacobi", "tol": 1e-8,
                   "note": "CPU restatement (port) of the reference path: dolfinx/PETSc are not installable here and the "
                           "reference checkout does not exist on the GPU box, so its own conjugate_gradient_solver.py is "
                           "restated in C/OpenMP (oracle/oracle_c.c) with Jacobi preconditioning"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count(),
                         "assembly_elements_per_s": r["asm_elems_per_s"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))



# --------------------------------------------------------------------------- parity / config 5 helpers
def rel_max(a, b):
    import torch
    return float((a - b).abs().max() / b.abs().max())


def parity_single(ctx, fem, fixed, g, f):
    """N = 1: (i) assembled vs matrix-free at full size, tol 1e-12; (ii) the CUDA path vs the CPU oracle's sparse
    direct solve on a BCC 6^3 (2 elements/strut) case of the same load -- oracle/ used as the checker only."""
    import torch
    from pylatticedso_b200 import lib as L, mesh as M
    from pylatticedso_b200.fem import BeamFEM
    from oracle import lattice_oracle as orc
    fem.build_pattern()        # the secondary sections replaced the resident pattern
    fem.vals = None
    ua, Ra, ia = fem.solve(fixed, g, f, tol=1e-12, maxiter=400000, precond=L.PC_BLOCK6)
    ua, Ra = ua.clone(), Ra.clone()
    um, Rm, im = fem.solve_matrix_free(fixed, g, f, tol=1e-12, maxiter=400000, precond=L.PC_BLOCK6)
    out = {"matrix_free_vs_assembled_u_rel": rel_max(um, ua), "matrix_free_vs_assembled_R_rel": rel_max(Rm, Ra),
           "true_relres_assembled": ia["true_relres"], "true_relres_matrix_free": im["true_relres"]}
    lat = M.synthetic_lattice("BCC", (6, 6, 6), [0.05])
    m = M.mesh_from_synthetic(lat, 2)
    fx, gg, ff = M.compression_bc(m)
    small = BeamFEM(m, E_MOD, NU, KAPPA, ctx=ctx)
    u, R, info = small.solve(fx, gg, ff, tol=1e-12, maxiter=100000, precond=L.PC_BLOCK6)
    grp = m.cell_of_elem.astype(np.int32)
    ng = int(grp.max()) + 1
    gr = small.compliance_gradient(u, grp, ng).cpu().numpy()
    en = np.stack([m.en0, m.en1], 1)
    K = orc.assemble_csr(m.xyz, en, m.rad, E_MOD, NU)
    uo, Ro = orc.solve_static(K, fx.astype(bool), gg, ff)
    go = orc.compliance_gradient(m.xyz, en, m.rad, uo, grp, ng, E_MOD, NU)
    out.update(u_rel=float(np.abs(u.cpu().numpy() - uo).max() / np.abs(uo).max()),
               R_rel=float(np.abs(R.cpu().numpy() - Ro).max() / np.abs(Ro).max()),
               grad_rel=float(np.abs(gr - go).max() / np.abs(go).max()),
               against="CPU oracle direct solve, BCC 6x6x6 m=2 (%d DOF); full-size figures are assembled vs matrix-free" % m.n_dof)
    out["ok"] = bool(out["u_rel"] < 1e-8 and out["R_rel"] < 1e-8 and out["grad_rel"] < 1e-6 and
                     out["matrix_free_vs_assembled_u_rel"] < 1e-8 and out["matrix_free_vs_assembled_R_rel"] < 1e-8)
    return out


def parity_sharded(ctx, dfem, mesh, fixed, g, f, rank):
    """N > 1: the sharded solve (peer-memory exchange when enabled) at tol 1e-12, assembled and matrix-free, against
    the same global system solved on rank 0's GPU alone; all-gathered owned entries are compared on rank 0."""
    import torch
    import torch.distributed as dist
    from pylatticedso_b200 import lib as L
    from pylatticedso_b200.fem import BeamFEM
    grp_g = mesh.cell_of_elem.astype(np.int32)
    ng = int(grp_g.max()) + 1
    res = {}
    sols = {}
    for name, solve in (("assembled", dfem.solve),
                        ("assembled_three_kernel", lambda **kw: dfem.solve(persistent=False, **kw)),
                        ("matrix_free", dfem.solve_matrix_free)):
        u, R, info = solve(tol=1e-12, maxiter=400000, precond=L.PC_BLOCK6)
        gr = dfem.compliance_gradient(u, grp_g, ng).cpu().numpy()
        sols[name] = (dfem.gather_owned(u), dfem.gather_owned(R), gr, info)
    if rank == 0:
        one = BeamFEM(mesh, E_MOD, NU, KAPPA, ctx=ctx)
        u0, R0, i0 = one.solve(fixed, g, f, tol=1e-12, maxiter=400000, precond=L.PC_BLOCK6)
        g0 = one.compliance_gradient(u0, grp_g, ng).cpu().numpy()
        u0, R0 = u0.cpu().numpy(), R0.cpu().numpy()
        ok = True
        for name, (ug, Rg, gr, info) in sols.items():
            e = dict(u_rel=float(np.abs(ug - u0).max() / np.abs(u0).max()), R_rel=float(np.abs(Rg - R0).max() / np.abs(R0).max()),
                     grad_rel=float(np.abs(gr - g0).max() / np.abs(g0).max()), iters=info["iters"], info=info["info"],
                     true_relres=info["true_relres"], cuda_graph=info.get("graph"), persistent_kernel=info.get("persistent", False))
            ok = ok and e["u_rel"] < 1e-8 and e["R_rel"] < 1e-8 and e["grad_rel"] < 1e-6
            res[name] = e
        res["against"] = "the same global system solved on rank 0 alone (single-GPU path), tol 1e-12, iters %d" % i0["iters"]
        res["u_rel"] = max(res[k]["u_rel"] for k in sols)
        res["R_rel"] = max(res[k]["R_rel"] for k in sols)
        res["grad_rel"] = max(res[k]["grad_rel"] for k in sols)
        res["ok"] = bool(ok)
        res["_u_full"] = u0             # popped by the joint-only section (not part of the JSON line)
        del one
    dist.barrier()
    return res


def config5_section(ctx, rank, world, hbm_peak, n=100):
    """BASELINE configs[4]: Octet n^3 (n = 100: 24 361 806 DOF), r = 0.03, 1 element per strut, uniaxial compression,
    block-Jacobi PCG to 1e-8, strong scaling over `world` GPUs.  Every rank generates ONLY its own cell layers."""
    import resource
    import torch
    import torch.distributed as dist
    from pylatticedso_b200 import lib as L
    from pylatticedso_b200 import distributed as D
    dev = ctx.device
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    if world == 1:
        from pylatticedso_b200.fem import BeamFEM
        lm, part = D.generate_slab("Octet", (n, n, n), [0.03], 1, 0, 1, device=dev)
        fixed, g, f = D.compression_bc_local(lm)

        class _Single:          # the single-GPU path behind the same few calls the sharded one offers
            def __init__(self):
                self.fem = BeamFEM(lm, E_MOD, NU, KAPPA, ctx=ctx)
                self.fem.build_pattern()
                self.n_owned, self.nnzb_owned = lm.n_nodes, self.fem.nnzb
                self.n_dof_global, self.n_elem_global = lm.n_dof, lm.n_elems
                self.vals = None
                t = lambda a_, d: torch.from_numpy(np.ascontiguousarray(a_, dtype=d)).to(dev)
                self.bc = (t(fixed, np.uint8), t(g, np.float64), t(f, np.float64))

            def assemble(self):
                self.fem.assemble()

            def solve(self, **kw):
                return self.fem.solve(*self.bc, keep_unconstrained=True, **kw)

            def solve_matrix_free(self, **kw):
                self.fem.vals = self.fem.vals_bc = None
                torch.cuda.empty_cache()
                return self.fem.solve_matrix_free(*self.bc, **kw)
        dfem = _Single()
    else:
        dfem = D.DistributedFEM.from_generator(ctx, "Octet", (n, n, n), [0.03], 1, E_MOD, NU, rank, world, KAPPA)
        fixed, g, f = D.compression_bc_local(dfem.lmesh)
        dfem.set_bc_local(fixed, g, f)
    t_gen_upload = time.perf_counter() - t0
    exch = "none"
    if world > 1:
        exch = "nccl"
        try:
            dfem.enable_p2p()
            exch = "nvlink-peer-memory"
        except Exception as e_:
            if rank == 0:
                print(f"bench.py: config5 peer-memory path unavailable ({e_}); using NCCL", file=sys.stderr)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b_, c = ev(), ev(), ev()
    a.record()
    dfem.assemble()
    b_.record()
    u, R, info = dfem.solve(tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, profile_iters=32)
    c.record()
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0 - 1e-3 * info["solve_ms"]
    asm_ms, tot_ms = a.elapsed_time(b_), a.elapsed_time(c)
    ua = u.clone()
    no = 6 * dfem.n_owned
    chk = torch.stack([ua[:no].abs().sum(), (R[:no] * ua[:no]).sum()])
    dfem.vals = None
    del R
    torch.cuda.empty_cache()
    um, Rm, im = dfem.solve_matrix_free(tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, want_reactions=False)
    diff = torch.stack([(um[:no] - ua[:no]).abs().max(), ua[:no].abs().max()])
    # the same matrix-free solve with the two-level preconditioner (block-Jacobi + rigid-body-mode coarse space,
    # csrc/coarse.cuh); set-up = temporary assembly + Galerkin product (+ all-reduce of the coarse matrix) + dense
    # factorisation, timed warm (second build).  A failure here must not take the headline down: it is reported instead.
    two = None
    try:
        for _ in range(2):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t1 = time.perf_counter()
            tl = dfem.fem.two_level(dfem.bc[0]) if world == 1 else dfem.two_level()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            tl_setup_ms = 1e3 * (time.perf_counter() - t1)
        if world == 1:
            u2, _, i2 = dfem.fem.solve_matrix_free(*dfem.bc, tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, want_reactions=False,
                                                   two_level=tl)
        else:
            u2, _, i2 = dfem.solve_matrix_free(tol=1e-8, maxiter=100000, precond=L.PC_BLOCK6, want_reactions=False, two_level=tl)
        d2 = torch.stack([(u2[:no] - um[:no]).abs().max(), um[:no].abs().max()])
        t2 = torch.tensor([i2["solve_ms"], tl_setup_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(d2, op=dist.ReduceOp.MAX)
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        two = {"solve_ms": float(t2[0]), "iters": i2["iters"], "info": i2["info"], "true_relres": i2["true_relres"],
               "setup_ms": float(t2[1]), "n_aggregates": tl.n_agg, "coarse_dof": 6 * tl.n_agg,
               "u_rel_vs_block_jacobi": float(d2[0] / d2[1]),
               "speedup_vs_block_jacobi": im["solve_ms"] / float(t2[0]),
               "note": "matrix-free PCG, M^-1 = D^-1 + Z E^+ Z^T; N > 1: coarse residual all-reduced once per iteration, separate "
                       "halo kernel; the solution is the same to the solver tolerance"}
        del tl, u2
        torch.cuda.empty_cache()
    except Exception as e_:      # noqa: BLE001
        two = {"error": f"{type(e_).__name__}: {e_}"[:300]}
        if rank == 0:
            print(f"bench.py: config5 two-level section failed: {e_}", file=sys.stderr)
    tt = torch.tensor([asm_ms, info["solve_ms"], im["solve_ms"], t_gen_upload, t_first,
                       resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6], dtype=torch.float64, device=dev)
    cnt = torch.tensor([dfem.n_owned, dfem.nnzb_owned], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        dist.all_reduce(chk)
        dist.all_reduce(cnt)
        ctx.p2p_destroy()
    nn, nz = int(cnt[0]), int(cnt[1])
    it_bytes = iteration_bytes(nn, nz, True)
    out = {
        "workload": f"Octet {n}x{n}x{n}, r=0.03, 1 element/strut, uniaxial compression, block-Jacobi PCG to 1e-8, strong scaling over {world} GPU(s)",
        "n_dof": dfem.n_dof_global, "n_elements": dfem.n_elem_global, "nnzb": nz, "exchange": exch,
        "assembled": {"solve_ms": float(tt[1]), "iters": info["iters"], "info": info["info"], "true_relres": info["true_relres"],
                      "dof_iters_per_s": dfem.n_dof_global * info["iters"] / (float(tt[1]) * 1e-3),
                      "assemble_ms": float(tt[0]), "assembly_elements_per_s": dfem.n_elem_global / (float(tt[0]) * 1e-3),
                      "iteration_frac_of_hbm": it_bytes * info["iters"] / (float(tt[1]) * 1e-3) / 1e9 / (hbm_peak * world),
                      "k_cg_spmv_frac_rank0": (spmv_bytes(dfem.n_owned, dfem.nnzb_owned) / (info["spmv_ms"] * 1e-3) / 1e9 / hbm_peak)
                      if info.get("spmv_ms", 0) > 0 else None,
                      "cuda_graph": info.get("graph")},
        "matrix_free": {"solve_ms": float(tt[2]), "iters": im["iters"], "info": im["info"], "true_relres": im["true_relres"],
                        "dof_iters_per_s": dfem.n_dof_global * im["iters"] / (float(tt[2]) * 1e-3),
                        "u_rel_vs_assembled": float(diff[0] / diff[1])},
        "matrix_free_two_level": two,
        "checksums": {"sum_abs_u": float(chk[0]), "u_dot_R": float(chk[1]),
                      "note": "all-reduced over ranks; equal across N to the solver tolerance"},
        "host": {"generate_upload_pattern_s_max": float(tt[3]), "time_to_first_iteration_s_max": float(tt[4]),
                 "peak_rss_gb_max": float(tt[5]),
                 "note": "every rank generates only its own cell layers + 1 overlap layer (distributed.generate_slab)"},
    }
    del dfem, u, ua, um
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from pylatticedso_b200 import lib as L
    from pylatticedso_b200.fem import BeamFEM

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.build()
    ctx = L.Context(local)
    dev = ctx.device
    hbm_peak, peak_src = measured_peaks()

    distributed = world > 1
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    lat, mesh, fixed, g, f = build_workload(world)
    n_dof_global, n_elem_global = mesh.n_dof, mesh.n_elems
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

    def ev():
        return torch.cuda.Event(enable_timing=True)

    if distributed:
        from pylatticedso_b200 import distributed as D
        ctx.comm_create(rank, world)
        dfem = D.DistributedFEM(ctx, mesh, E_MOD, NU, rank, world, KAPPA)
        dfem.set_bc(fixed, g, f)
        comm_mode = "nccl"
        if not args.nccl:
            try:
                dfem.enable_p2p()
                comm_mode = "nvlink-peer-memory"
            except Exception as e_:  # still a GPU path: NCCL send/recv + all-reduce
                if rank == 0:
                    print(f"bench.py: peer-memory path unavailable ({e_}); using NCCL", file=sys.stderr)
        lm = dfem.lmesh
        host = {k: pin(v) for k, v in dict(x=lm.x, y=lm.y, z=lm.z, en0=lm.en0, en1=lm.en1, rad=lm.rad,
                                           fixed=fixed[dfem.dofs], g=g[dfem.dofs], f=f[dfem.dofs]).items()}
        vals = torch.empty(dfem.nnzb * 36, dtype=torch.float64, device=dev)
        vals_bc = torch.empty_like(vals)
        b_d = torch.empty(6 * dfem.n_local, dtype=torch.float64, device=dev)
        u_d = torch.empty_like(b_d)
        n_nodes, nnzb = dfem.n_owned, dfem.nnzb_owned

        def step(profile):
            e = [ev() for _ in range(2)]
            e[0].record()
            dfem.assemble(out=vals)
            e[1].record()
            u, R, info = dfem.solve(tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6, vals_bc=vals_bc, b=b_d, u=u_d,
                                    profile_iters=profile)
            e.append(ev()); e[2].record()
            return [e[0], e[1], e[1], e[2]], info, (u, R)

        def step_mf(profile):
            e = [ev() for _ in range(2)]
            e[0].record()
            u, R, info = dfem.solve_matrix_free(tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6, b=b_d, u=u_d,
                                                profile_iters=profile)
            e[1].record()
            return e, info

        dev_targets = (("x", dfem.x), ("y", dfem.y), ("z", dfem.z), ("en0", dfem.en0), ("en1", dfem.en1),
                       ("rad", dfem.rad), ("fixed", dfem.fixed_d), ("g", dfem.g_d), ("f", dfem.f_d))
        n_out = 6 * dfem.n_local
        pattern_ms = None
    else:
        host = {k: pin(v) for k, v in dict(x=mesh.x, y=mesh.y, z=mesh.z, en0=mesh.en0, en1=mesh.en1, rad=mesh.rad,
                                           fixed=fixed, g=g, f=f).items()}
        fem = BeamFEM(mesh, E_MOD, NU, KAPPA, ctx=ctx)
        t0 = ev(); t1 = ev()
        t0.record(); fem.build_pattern(); t1.record(); torch.cuda.synchronize()
        pattern_ms = t0.elapsed_time(t1)
        fixed_d, g_d, f_d = (host[k].to(dev) for k in ("fixed", "g", "f"))
        vals = torch.empty(fem.nnzb * 36, dtype=torch.float64, device=dev)
        vals_bc = torch.empty_like(vals)
        n_nodes, nnzb = fem.n_nodes, fem.nnzb
        b_d = torch.empty(fem.n_dof, dtype=torch.float64, device=dev)
        u_d = torch.empty_like(b_d)
        R_d = torch.empty_like(b_d)

        def step(profile):
            """One pass of the hot path with device-resident inputs. Returns per-phase events + PCG info."""
            e = [ev() for _ in range(4)]
            e[0].record()
            ctx.assemble_bsr(fem.x, fem.y, fem.z, fem.en0, fem.en1, fem.rad, n_nodes, nnzb, E_MOD, NU, KAPPA, out=vals)
            e[1].record()
            ctx.check(ctx.lib.lat_apply_dirichlet(ctx.h, L._ptr(fem.rowptr), L._ptr(fem.colidx), n_nodes, L._ptr(vals),
                                                  L._ptr(fixed_d), L._ptr(g_d), L._ptr(f_d), L._ptr(vals_bc), L._ptr(b_d)))
            e[2].record()
            u, info = ctx.pcg(fem.rowptr, fem.colidx, vals_bc, b_d, x=u_d, tol=1e-8, maxiter=200000,
                              precond=L.PC_BLOCK6, profile_iters=profile)
            ctx.set_dirichlet_values(fixed_d, g_d, u)
            ctx.spmv(fem.rowptr, fem.colidx, vals, u, out=R_d)
            e[3].record()
            return e, info, (u_d, R_d)

        def step_mf(profile):
            """The same system without an assembled matrix (operator set-up, lifting, PCG, reactions)."""
            e = [ev() for _ in range(2)]
            e[0].record()
            ctx.matfree_setup(fem.x, fem.y, fem.z, fem.en0, fem.en1, fem.rad, n_nodes, E_MOD, NU, KAPPA, fixed=fixed_d)
            ctx.matfree_rhs(g_d, f_d, out=b_d)
            u, info = ctx.pcg_matfree(b_d, x=u_d, tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6, profile_iters=profile)
            ctx.set_dirichlet_values(fixed_d, g_d, u)
            ctx.matfree_apply(u, out=R_d, eliminated=False)
            e[1].record()
            return e, info

        def step_condensed():
            """Same load case through the exact joint-only system (struts condensed; BeamFEM.solve_condensed)."""
            e = [ev() for _ in range(2)]
            e[0].record()
            u, R, info = fem.solve_condensed(fixed, g, f, tol=1e-8, precond=L.PC_BLOCK6)
            e[1].record()
            return e, info

        dev_targets = (("x", fem.x), ("y", fem.y), ("z", fem.z), ("en0", fem.en0), ("en1", fem.en1), ("rad", fem.rad),
                       ("fixed", fixed_d), ("g", g_d), ("f", f_d))
        n_out = fem.n_dof

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(0)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    tot_ms = asm_ms = solve_ms = 0.0
    iters = 0
    spmv_ms, upd_ms, nprof = [], [], 0
    for _ in range(args.steps):
        flush.fill_(1.0)            # evict the previous step's matrix from L2 (untimed)
        barrier()
        # the solve is ONE launch of the persistent on-chip kernel per GPU (N > 1: halo exchange and all-reduce inside its
        # grid barriers), timed by the library with CUDA events on its stream (info["solve_ms"])
        e, info, _ = step(0)
        barrier()
        persistent = bool(info.get("persistent", False))
        tot_ms += e[0].elapsed_time(e[3])
        asm_ms += e[0].elapsed_time(e[1])
        solve_ms += info["solve_ms"]
        iters += info["iters"]
        assert info["info"] == 0, f"PCG did not converge: {info}"
        spmv_ms.append(info.get("spmv_ms", 0.0)); upd_ms.append(info.get("update_ms", 0.0)); nprof += info.get("profiled", 0)
    launches = ctx.launches - launches0
    clocks = sampler.stop()
    # ---- end-to-end through the public array-level API, inputs in pinned HOST memory, everything inside the timed
    #      region: H2D of mesh + BCs, device allocation, pattern build, assembly, elimination, PCG, reactions, D2H.
    #      N = 1: BeamFEM(mesh) + BeamFEM.solve(fixed, g, f);  N > 1: DistributedFEM.upload() + set_bc_local + solve.
    from pylatticedso_b200.mesh import BeamMesh

    def pinned_mesh(m_):
        arrs = {k: pin(getattr(m_, k)) for k in ("x", "y", "z", "en0", "en1", "rad")}
        pm = BeamMesh(**{k: v.numpy() for k, v in arrs.items()}, beam_of_elem=m_.beam_of_elem, chain=m_.chain,
                      n_points=m_.n_points, point_index=m_.point_index, cell_of_elem=m_.cell_of_elem, meta=dict(m_.meta))
        return pm, arrs

    if distributed:
        pm, keep_pinned = pinned_mesh(dfem.lmesh)
        bc_pin = [pin(v) for v in (fixed[dfem.dofs], g[dfem.dofs], f[dfem.dofs])]
    else:
        pm, keep_pinned = pinned_mesh(mesh)
        bc_pin = [pin(v) for v in (fixed, g, f)]
    h2d = sum(t.numel() * t.element_size() for t in list(keep_pinned.values()) + bc_pin)
    u_host = torch.empty(n_out, dtype=torch.float64).pin_memory()
    R_host = torch.empty(n_out, dtype=torch.float64).pin_memory()
    d2h = 2 * n_out * 8
    e2e_ms, e2e_iters = 0.0, 0
    e2e_warm = max(2, args.warmup)      # the first passes grow the allocator's pools (181 / 115 / 21.8 / 21.8 ms)
    for s_ in range(e2e_warm + args.steps):
        flush.fill_(1.0)
        barrier()
        a, b_ = ev(), ev()
        a.record()
        if distributed:
            dfem.lmesh = pm
            dfem.upload()
            dfem.set_bc_local(*[t.numpy() for t in bc_pin])
            uu, RR, info = dfem.solve(tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6)
        else:
            fem_e = BeamFEM(pm, E_MOD, NU, KAPPA, ctx=ctx)
            uu, RR, info = fem_e.solve(bc_pin[0].numpy(), bc_pin[1].numpy(), bc_pin[2].numpy(), tol=1e-8,
                                       maxiter=200000, precond=L.PC_BLOCK6)
        u_host.copy_(uu, non_blocking=True)
        R_host.copy_(RR, non_blocking=True)
        b_.record()
        barrier()
        assert info["info"] == 0, f"e2e PCG did not converge: {info}"
        if args.verbose and rank == 0:
            print(f"bench.py: e2e pass {s_}: {a.elapsed_time(b_):.2f} ms, solve {info['solve_ms']:.2f} ms, {info['iters']} iterations", file=sys.stderr)
        if s_ >= e2e_warm:
            e2e_ms += a.elapsed_time(b_)
            e2e_iters += info["iters"]
        if not distributed:
            del fem_e
    if not distributed:       # the e2e passes rebuilt the resident pattern; make the bench operator current again
        fem.build_pattern()
    # ---- secondary: the matrix-free operator on the same workload (same barriers, L2 flush, CUDA events)
    mf_ms, mf_iters, mf_prod = 0.0, 0, []
    for s_ in range(args.warmup + args.steps):
        flush.fill_(1.0)
        barrier()
        e, info = step_mf(32 if s_ >= args.warmup else 0)
        barrier()
        assert info["info"] == 0, f"matrix-free PCG did not converge: {info}"
        if s_ >= args.warmup:
            mf_ms += e[0].elapsed_time(e[1])
            mf_iters += info["iters"]
            mf_prod.append(info.get("spmv_ms", 0.0))
    # ---- secondary (N = 1): the same solve through the three-kernel iteration (what larger systems use), with
    #      CUDA events around k_cg_spmv for the first 32 iterations
    three = None
    if True:
        t_ms, t_it, t_sp, t_up = 0.0, 0, [], []
        for s_ in range(1 + args.steps):
            flush.fill_(1.0)
            barrier()
            if distributed:
                _, _, info3 = dfem.solve(tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6, vals_bc=vals_bc, b=b_d, u=u_d,
                                         profile_iters=32, persistent=False)
            else:
                _, info3 = ctx.pcg(fem.rowptr, fem.colidx, vals_bc, b_d, x=u_d, tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6,
                                   profile_iters=32, persistent=False)
            barrier()
            if s_ >= 1:
                t_ms += info3["solve_ms"]; t_it += info3["iters"]; t_sp.append(info3["spmv_ms"]); t_up.append(info3["update_ms"])
        three = dict(ms=t_ms / args.steps, iters=t_it / args.steps, spmv_ms=float(np.mean(t_sp)), update_ms=float(np.mean(t_up)))
    # ---- secondary (N = 1): the strut-condensed joint-only solve; LAST, because it replaces the resident pattern
    cond = None
    if not distributed:
        cms, cit = 0.0, 0
        for s_ in range(args.warmup + args.steps):
            flush.fill_(1.0)
            barrier()
            e, info = step_condensed()
            barrier()
            assert info["info"] == 0, f"condensed PCG did not converge: {info}"
            if s_ >= args.warmup:
                cms += e[0].elapsed_time(e[1]); cit += info["iters"]
        cond = dict(ms=cms / args.steps, iters=cit / args.steps, n_dof=info["n_dof_condensed"])
    # ---- secondary (N = 1): the other two rates SURVEY section 8(d) names -- gradient elements/s and Schur cells/s
    extra = None
    if not distributed:
        from pylatticedso_b200.schur import synthetic_cell_batch
        grp = torch.from_numpy(np.ascontiguousarray(mesh.cell_of_elem, dtype=np.int32)).to(dev)
        ng = int(mesh.cell_of_elem.max()) + 1
        gfun = lambda: ctx.compliance_grad(fem.x, fem.y, fem.z, fem.en0, fem.en1, fem.rad, grp, ng, u_d, E_MOD, NU, KAPPA)
        for _ in range(3):
            gfun()
        a, b_ = ev(), ev()
        a.record()
        for _ in range(10):
            gfun()
        b_.record(); torch.cuda.synchronize()
        g_ms = a.elapsed_time(b_) / 10
        n_sc = 50000
        batch, _bnd = synthetic_cell_batch(ctx, "BCC", 0.02 + 0.06 * np.random.default_rng(1).random(n_sc), 18, E_MOD, NU)
        batch.schur()
        a, b_ = ev(), ev()
        a.record(); batch.schur(); b_.record(); torch.cuda.synchronize()
        s_ms = a.elapsed_time(b_)
        extra = dict(grad_eps=mesh.n_elems / (g_ms * 1e-3), grad_ms=g_ms, schur_cps=n_sc / (s_ms * 1e-3), schur_ms=s_ms, schur_cells=n_sc)
        del batch
        # N4: surrogate Schur complements of BASELINE config 4's 216 000 cells from the reference's stored BCC basis
        from pylatticedso_b200 import surrogate
        rb = np.load(os.path.join(ROOT, "tests", "golden", "reduced_basis_BCC_tol_1e-6.npz"))
        sur = surrogate.SchurSurrogate(rb, "RBF", ctx=ctx)
        xq = surrogate._dev(ctx, np.random.default_rng(44).uniform(0.01, 0.1, (216000, 1)))
        s_out = torch.empty((216000, 48, 48), dtype=torch.float64, device=dev)
        sur.expand_device(sur.alphas_device(xq), out=s_out)
        a, b_ = ev(), ev()
        a.record(); sur.expand_device(sur.alphas_device(xq), out=s_out); b_.record(); torch.cuda.synchronize()
        extra["surrogate_ms"] = a.elapsed_time(b_)
        del s_out, sur
        torch.cuda.empty_cache()
        # BASELINE configs[3] end to end through the DDM path, against the full FEM solve of the same lattice
        import importlib.util
        spec = importlib.util.spec_from_file_location("ddm_config3", os.path.join(ROOT, "tools", "ddm_config3.py"))
        ddm3 = importlib.util.module_from_spec(spec); spec.loader.exec_module(ddm3)
        extra["ddm3"] = ddm3.run(ctx, 60, 1, tol=1e-10, verbose=False)
        torch.cuda.empty_cache()
    # ---- parity block (driver-visible): see parity_single / parity_sharded
    if distributed:
        parity = parity_sharded(ctx, dfem, mesh, fixed, g, f, rank)
    else:
        parity = parity_single(ctx, fem, fixed, g, f)
    # ---- secondary (N > 1): the strut-condensed joint-only system, sharded (DistributedJointFEM), same load case;
    #      after the parity block because it needs the context's peer-memory arena for its own (smaller) vectors
    if distributed:
        if getattr(dfem, "p2p", False):
            ctx.p2p_destroy()
            dfem.p2p = False
        jf = D.DistributedJointFEM(ctx, mesh, E_MOD, NU, rank, world, KAPPA)
        nj = 6 * mesh.n_points
        jf.set_bc(fixed[:nj], g[:nj], f[:nj])
        if comm_mode == "nvlink-peer-memory":
            jf.enable_p2p()
        cms, cit = 0.0, 0
        for s_ in range(args.warmup + args.steps):
            flush.fill_(1.0)
            barrier()
            a, b_ = ev(), ev()
            a.record()
            jf.assemble()
            uj, Rj, info = jf.solve(tol=1e-8, maxiter=200000, precond=L.PC_BLOCK6)
            b_.record()
            barrier()
            assert info["info"] == 0, f"joint-only PCG did not converge: {info}"
            if s_ >= args.warmup:
                cms += a.elapsed_time(b_); cit += info["iters"]
        # parity of the sharded joint-only solve: against the sharded FULL solve of the parity block at the lattice points
        uj12, _, _ = jf.solve(tol=1e-12, maxiter=400000, precond=L.PC_BLOCK6)
        ujg = jf.gather_owned(uj12)
        cond = dict(ms=cms / args.steps, iters=cit / args.steps, n_dof=nj, persistent=bool(info.get("persistent", False)))
        if rank == 0 and parity.get("_u_full") is not None:
            u_full = parity.pop("_u_full")
            cond["u_rel_vs_full_solve"] = float(np.abs(ujg - u_full[:nj]).max() / np.abs(u_full).max())
            parity["joint_only_u_rel"] = cond["u_rel_vs_full_solve"]
            parity["ok"] = bool(parity["ok"] and cond["u_rel_vs_full_solve"] < 1e-8)
        parity.pop("_u_full", None)
        if getattr(jf, "p2p", False):
            ctx.p2p_destroy()
        del jf
    # ---- BASELINE configs[4] on the same N GPUs
    cfg5 = None
    if not args.no_config5:
        if distributed and getattr(dfem, "p2p", False):
            ctx.p2p_destroy()
        flush = None
        torch.cuda.empty_cache()
        cfg5 = config5_section(ctx, rank, world, hbm_peak, n=args.config5_n)
    res = dict(extra=extra, cond=cond, three=three, persistent=persistent, mf_ms=mf_ms, mf_iters=mf_iters, mf_prod_ms=float(np.mean(mf_prod)),
               tot_ms=tot_ms, asm_ms=asm_ms, solve_ms=solve_ms, iters=iters, launches=launches, clocks=clocks,
               spmv_ms=float(np.mean(spmv_ms)), update_ms=float(np.mean(upd_ms)), nprof=nprof,
               e2e_ms=e2e_ms, e2e_iters=e2e_iters, h2d=h2d, d2h=d2h, pattern_ms=pattern_ms,
               n_nodes=n_nodes, nnzb=nnzb)

    # max over ranks of the timed region
    tot_ms = res["tot_ms"]
    if world > 1:
        t = torch.tensor([res["tot_ms"], res["e2e_ms"], res["asm_ms"], res["solve_ms"], res["mf_ms"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_ms, res["e2e_ms"], res["asm_ms"], res["solve_ms"], res["mf_ms"] = (float(v) for v in t)
        if res.get("cond"):
            tc = torch.tensor([res["cond"]["ms"]], dtype=torch.float64, device=dev)
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            res["cond"]["ms"] = float(tc[0])
        t2 = torch.tensor([res["h2d"], res["d2h"], res["launches"], res["n_nodes"], res["nnzb"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        res["h2d"], res["d2h"], res["launches"] = int(t2[0]), int(t2[1]), int(t2[2])
        res["n_nodes_all"], res["nnzb_all"] = int(t2[3]), int(t2[4])
        ctx.comm_destroy()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = n_dof_global * res["iters"] / (tot_ms * 1e-3)
    e2e_val = n_dof_global * res["e2e_iters"] / (res["e2e_ms"] * 1e-3)
    nn, nz = res["n_nodes"], res["nnzb"]
    # rank 0's own kernel: its launch processes rank 0's rows (nn, nz are rank-local for N > 1)
    ach = spmv_bytes(nn, nz) / (res["spmv_ms"] * 1e-3) / 1e9 if res["spmv_ms"] > 0 else None
    it_bytes = iteration_bytes(nn, nz, True) if world == 1 else iteration_bytes(res["n_nodes_all"], res["nnzb_all"], True)
    it_gbs = it_bytes * res["iters"] / (res["solve_ms"] * 1e-3) / 1e9
    if res["persistent"]:
        # dominant kernel = the whole solve: ONE launch of k_pcg_persist per step.  Algorithmic bytes per launch =
        # SURVEY 8(d)'s iteration figure x iterations; the kernel keeps r, p, s, w on chip, so what it really moves per
        # iteration is the matrix + the preconditioner rows + u and x (design_bytes_per_iteration).
        it_b = iteration_bytes(nn, nz, True)
        design_b = nz * 292 + nn * 4 + 6 * nn * (28 + 4 * 8)     # matrix+index, packed inverse blocks, u w/r + x r/w
        per_step_iters = res["iters"] / args.steps
        launch_ms = res["solve_ms"] / args.steps
        ach = it_b * per_step_iters / (launch_ms * 1e-3) / 1e9
        roofline = {"kernel": "k_pcg_persist (the whole block-Jacobi PCG solve as one persistent cooperative kernel: BSR 6x6 product, "
                              "dot products, grid reductions, vector updates and preconditioner; r, p, s, w in shared memory)",
                    "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": NCU_TRAFFIC_PERSIST_PER_ITER * per_step_iters if (NCU_TRAFFIC_PERSIST_PER_ITER and world == 1) else None,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full, scaled to this "
                                      "step's iteration count (profiles/r02_ncu_persist.txt)",
                    "peak_source": peak_src, "bytes_per_launch": it_b * per_step_iters,
                    "bytes_per_iteration_survey_8d": it_b, "design_bytes_per_iteration": design_b,
                    "frac_on_design_bytes": design_b * per_step_iters / (launch_ms * 1e-3) / 1e9 / hbm_peak,
                    "avg_launch_ms": launch_ms, "us_per_iteration": 1e3 * launch_ms / per_step_iters,
                    "launches_timed": args.steps,
                    "scope": None if world == 1 else "rank 0's kernel over rank 0's slab against ONE GPU's peak; the halo exchange and the "
                                                     "rank-level all-reduce happen inside the kernel's two grid barriers",
                    "dram_frac": (NCU_TRAFFIC_PERSIST_PER_ITER * per_step_iters / (launch_ms * 1e-3) / 1e9 / hbm_peak) if world == 1 else None,
                    "note": "frac uses SURVEY 8(d)'s algorithmic bytes of a PCG iteration (SpMV + 96 B/DOF + 28 B/DOF) and can "
                            "exceed 1: the kernel never moves the 56 MB/iteration of vector traffic those include (r, p, s, w "
                            "live in shared memory) and ~45 % of the 98 MB matrix stays L2-resident between iterations "
                            "(evict_last policy, profiles/r02_persist_l2keep_ab.txt), so DRAM sees `traffic` = 40 MB per "
                            "iteration (dram_frac); the product phase then runs at the L2->SM limit (98 MB in 11.5 us) and "
                            "half of the iteration is the two grid barriers + the on-chip update"}
    else:
        roofline = {"kernel": "k_cg_spmv (BSR 6x6 SpMV w = A u fused with the partial sums of (r,u), (w,u), (r,r))", "bound": "hbm",
                    "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": (ach / hbm_peak) if ach else None,
                    "traffic": None, "peak_source": peak_src, "bytes_per_launch": spmv_bytes(nn, nz),
                    "avg_launch_ms": res["spmv_ms"], "launches_timed": res["nprof"],
                    "note": None if world == 1 else "rank 0's kernel over rank 0's slab (per-GPU peak); see pcg.iteration_frac_of_hbm "
                                                    "for the whole job against the aggregate peak"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(world),
                   "n_dof": n_dof_global, "n_elements": n_elem_global, "precond": "block-jacobi-6x6", "tol": 1e-8,
                   "iterations_per_step": res["iters"] / args.steps,
                   "l2": "L2 flushed (256 MB write) between steps; within a step the 98 MB matrix is re-streamed "
                         "every PCG iteration (working set ~ L2 size: part of it stays L2-resident, see roofline.traffic)",
                   "solver": "persistent on-chip kernel (csrc/pcg_persist.cuh)" if res["persistent"] else "three-kernel iteration in CUDA graphs",
                   "parallelism": "single" if world == 1 else f"slab{world}",
                   "exchange": None if world == 1 else comm_mode},
        "assembly": {"value": n_elem_global * args.steps / (res["asm_ms"] * 1e-3), "unit": "elements/s",
                     "ms": res["asm_ms"] / args.steps, "mode": "rows (deterministic, fused element generation, 256-bit stores)",
                     "pattern_build_ms_one_off": res.get("pattern_ms")},
        "pcg": {"solve_ms_per_step": res["solve_ms"] / args.steps, "iteration_GBps_survey_bytes": it_gbs,
                "iteration_frac_of_hbm": it_gbs / (hbm_peak * world), "update_kernel_ms": res["update_ms"]},
        "roofline": roofline,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                "ms_per_step": res["e2e_ms"] / args.steps,
                "path": ("BeamFEM(mesh) + BeamFEM.solve(fixed, g, f)" if world == 1 else
                         "DistributedFEM.upload() + set_bc_local + solve") + ": host arrays in pinned memory -> H2D, device "
                        "allocation, BSR pattern build, assembly, elimination, PCG, reactions, D2H of u and R, all timed"},
        "matrix_free": {"value": n_dof_global * res["mf_iters"] / (res["mf_ms"] * 1e-3), "unit": UNIT,
                        "ms_per_step": res["mf_ms"] / args.steps, "iterations_per_step": res["mf_iters"] / args.steps,
                        "product_kernel_ms": res["mf_prod_ms"],
                        "note": "same system, tolerance and preconditioner solved WITHOUT an assembled matrix "
                                "(csrc/matfree.cuh: the element action is regenerated from the geometry in every "
                                "product); not the headline because BASELINE's metric is quoted on the assembled "
                                "BSR path, timed after the headline region"},
        "gpu_launches": res["launches"],
        "clocks": res["clocks"],
        "parity": parity,
    }
    if cfg5 is not None:
        line["config5"] = cfg5
    if res.get("extra"):
        x_ = res["extra"]
        line["gradient"] = {"value": x_["grad_eps"], "unit": "elements/s", "ms": x_["grad_ms"],
                            "note": "lat_compliance_grad (adjoint compliance sensitivity, 168 B/element) on the bench mesh"}
        line["surrogate"] = {"value": 216000 / (x_["surrogate_ms"] * 1e-3), "unit": "cells/s", "ms": x_["surrogate_ms"], "cells": 216000,
                             "note": "N4: thin-plate-spline RBF coefficients + basis @ alphas on the FP64 tensor cores (DMMA), 48x48 "
                                     "Schur complements from the reference's stored BCC reduced basis (k = 5)"}
        d3 = x_["ddm3"]
        line["ddm_config3"] = {"workload": "BCC 60x60x60 = 216000 cells, per-cell radii (rng 44), 1 element per strut, compression",
                               "interface_dof": d3["interface_dof"], "condense_ms": d3["condense_ms"],
                               "interface_assembly_ms": d3["interface_assembly_ms"], "pcg_iters": d3["ddm_iters"],
                               "pcg_ms": d3["ddm_pcg_ms"], "full_fem_dof": d3["fem_dof"], "full_fem_pcg_ms": d3["fem_pcg_ms"],
                               "full_fem_iters": d3["fem_iters"], "u_rel_vs_full_fem": d3["u_rel"], "R_rel_vs_full_fem": d3["R_rel"],
                               "two_level": {"pcg_iters": d3["two_level_iters"], "pcg_ms": d3["two_level_pcg_ms"],
                                             "wall_ms_with_setup": d3["two_level_wall_ms"], "info": d3["two_level_info"],
                                             "u_rel_vs_block_jacobi": d3["two_level_u_rel"]},
                               "note": "BASELINE configs[3] end to end (tools/ddm_config3.py): batched condensation -> assembled "
                                       "interface operator -> block-Jacobi PCG, both solves to 1e-10; static condensation is exact, "
                                       "so the corner displacements must equal the full FEM solve"}
        parity["ddm_config3_u_rel"] = d3["u_rel"]
        parity["ok"] = bool(parity["ok"] and d3["u_rel"] < 1e-7 and d3["ddm_info"] == 0 and d3["two_level_info"] == 0
                            and d3["two_level_u_rel"] < 1e-7)
        line["schur"] = {"value": x_["schur_cps"], "unit": "cells/s", "ms": x_["schur_ms"], "cells": x_["schur_cells"],
                         "note": "lat_schur_batch_chains: BCC cells at the reference mesh density (18 elements per strut, "
                                 "870 DOF -> 48 boundary DOF), strut pre-pass + joint-only condensation"}
    if res.get("three"):
        t3 = res["three"]
        b3 = spmv_bytes(nn, nz)
        if world > 1:       # t3["ms"] is rank 0's library-timed solve; the ranks run in lock step
            b3 = spmv_bytes(nn, nz)
        line["three_kernel_path"] = {"solve_ms_per_step": t3["ms"], "iterations_per_step": t3["iters"],
                                     "value": n_dof_global * t3["iters"] / (t3["ms"] * 1e-3), "unit": UNIT,
                                     "k_cg_spmv_ms": t3["spmv_ms"], "k_cg_spmv_GBps": b3 / (t3["spmv_ms"] * 1e-3) / 1e9,
                                     "k_cg_spmv_frac_of_hbm": b3 / (t3["spmv_ms"] * 1e-3) / 1e9 / hbm_peak,
                                     "k_cg_update_ms": t3["update_ms"],
                                     "note": "the PCG of the headline through k_cg_update / k_cg_spmv / k_cg_reduce in CUDA graphs: the path "
                                             "systems too large for the on-chip kernel (config5) run; solve only, events by the library"}
    if res.get("cond"):
        c = res["cond"]
        line["strut_condensed"] = {"ms_per_step": c["ms"], "n_dof": c["n_dof"], "iterations_per_step": c["iters"],
                                   "speedup_vs_headline_step": (tot_ms / args.steps) / c["ms"],
                                   "sharded": world > 1, "persistent_kernel": c.get("persistent"),
                                   "u_rel_vs_full_solve": c.get("u_rel_vs_full_solve"),
                                   "note": "same load case and tolerance through the exact joint-only system (every strut "
                                           "condensed onto its two lattice points, lat_assemble_bsr_struts): identical "
                                           "lattice-point displacements and reactions; a time-to-solution figure, not "
                                           "comparable in DOF-iterations/s (fewer DOFs AND fewer iterations); includes the "
                                           "pattern build of the joint mesh and host-side set-up"}
    if world == 1 and not args.no_cpu_baseline:
        c = cpu_reference_run(1, 0, pcg_iters=3000)
        line["cpu_baseline"] = {"value": c["value"], "unit": UNIT, "cores": c["threads"], "kind": "port",
                                "sample": f"1 step: C/OpenMP (oracle/oracle_c.c, a port: the reference checkout is not on the box) assembly of the full mesh + Jacobi-PCG "
                                          f"to 1e-8 capped at 3000 iterations ({c['iters']} run), {c['threads']} threads",
                                "host_cores_available": os.cpu_count(),
                                "assembly_elements_per_s": c["asm_elems_per_s"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not parity.get("ok", False):
        print(f"bench.py: PARITY FAILURE {parity}", file=sys.stderr)
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl", action="store_true", help="multi-GPU: NCCL halo/all-reduce instead of NVLink peer memory")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the Octet 100^3 section (BASELINE configs[4])")
    ap.add_argument("--config5-n", type=int, default=100, help="cells per side of the config5 octet lattice")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
