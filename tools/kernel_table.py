"""Per-kernel throughput table on large configurations (CUDA events, 3 warm-up + 10 timed launches,
inputs larger than L2).  Writes profiles/r01_kernel_table.txt-style lines to stdout."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
from pylatticedso_b200.schur import synthetic_cell_batch
E, NU = 1013.0, 0.3
ctx = L.Context(); dev = ctx.device
peak = 6554.6
try: peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: pass
t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def line(name, cfg, units, unit_name, ms, bytes_):
    gbs = bytes_ / ms / 1e6
    print(f"{name:34s} {cfg:28s} {units/ms/1e6:10.2f} M{unit_name}/ms... {units/(ms*1e-3):12.4e} {unit_name}/s  {ms*1e3:10.1f} us  {gbs:8.0f} GB/s  frac={gbs/peak:5.2f}", flush=True)

for geom, n, m_, r in (("BCC", 60, 1, 0.05), ("Octet", 40, 1, 0.03)):
    lat = M.synthetic_lattice(geom, (n, n, n), [r]); mesh = M.mesh_from_synthetic(lat, m_)
    cfg = f"{geom} {n}^3 m={m_}"
    x, y, z, en0, en1, rad = t(mesh.x, np.float64), t(mesh.y, np.float64), t(mesh.z, np.float64), t(mesh.en0, np.int32), t(mesh.en1, np.int32), t(mesh.rad, np.float64)
    Ecount, N = mesh.n_elems, mesh.n_nodes
    ms = timeit(lambda: ctx.bsr_pattern(en0, en1, N), n=3, warm=1)
    rowptr, colidx = ctx.bsr_pattern(en0, en1, N); nnzb = colidx.numel()
    print(f"{'pattern build (one-off)':34s} {cfg:28s} {ms:10.2f} ms   N={N} E={Ecount} nnzb={nnzb}", flush=True)
    vals = torch.empty(nnzb * 36, dtype=torch.float64, device=dev)
    ms = timeit(lambda: ctx.assemble_bsr(x, y, z, en0, en1, rad, N, nnzb, E, NU, mode=L.ASM_GATHER, out=vals))
    line("k_assemble_gather (fused)", cfg, Ecount, "elements", ms, Ecount * 1216)
    print(f"{'':34s} {'  actual HBM write 288 B/block':28s} {nnzb*288/ms/1e6:8.0f} GB/s  frac={nnzb*288/ms/1e6/peak:5.2f}")
    ms = timeit(lambda: ctx.assemble_bsr(x, y, z, en0, en1, rad, N, nnzb, E, NU, mode=L.ASM_ROWS, out=vals))
    line("k_assemble_rows (fused)", cfg, Ecount, "elements", ms, Ecount * 1216)
    print(f"{'':34s} {'  actual HBM write 288 B/block':28s} {nnzb*288/ms/1e6:8.0f} GB/s  frac={nnzb*288/ms/1e6/peak:5.2f}")
    ms = timeit(lambda: ctx.assemble_bsr(x, y, z, en0, en1, rad, N, nnzb, E, NU, mode=L.ASM_ATOMIC, out=vals))
    line("k_assemble_atomic", cfg, Ecount, "elements", ms, Ecount * 1216)
    ctx.assemble_bsr(x, y, z, en0, en1, rad, N, nnzb, E, NU, out=vals)
    if Ecount * 1152 < 6e9:
        Ke = torch.empty((Ecount, 12, 12), dtype=torch.float64, device=dev)
        ms = timeit(lambda: ctx.check(ctx.lib.lat_elem_stiffness(ctx.h, L._ptr(x), L._ptr(y), L._ptr(z), L._ptr(en0), L._ptr(en1), L._ptr(rad), Ecount, E, NU, 0.9, 0, L._ptr(Ke))))
        line("k_elem_stiffness (K_e out)", cfg, Ecount, "elements", ms, Ecount * 1216)
        del Ke
    u = torch.randn(6 * N, dtype=torch.float64, device=dev); yv = torch.empty_like(u)
    ms = timeit(lambda: ctx.spmv(rowptr, colidx, vals, u, out=yv))
    line("k_bsr_spmv", cfg, 6 * N, "DOF", ms, nnzb * 292 + N * 100)
    grp = t(mesh.cell_of_elem, np.int32); ng = int(mesh.cell_of_elem.max()) + 1
    ms = timeit(lambda: ctx.compliance_grad(x, y, z, en0, en1, rad, grp, ng, u, E, NU))
    line("k_compliance_grad", cfg, Ecount, "elements", ms, Ecount * 168)
    fixed, g, f = M.compression_bc(mesh)
    fd, gd, fv = t(fixed, np.uint8), t(g, np.float64), t(f, np.float64)
    vbc = torch.empty_like(vals); b = torch.empty(6 * N, dtype=torch.float64, device=dev)
    ms = timeit(lambda: ctx.check(ctx.lib.lat_apply_dirichlet(ctx.h, L._ptr(rowptr), L._ptr(colidx), N, L._ptr(vals), L._ptr(fd), L._ptr(gd), L._ptr(fv), L._ptr(vbc), L._ptr(b))), n=5)
    line("lat_apply_dirichlet (4 kernels)", cfg, nnzb, "blocks", ms, nnzb * (576 + 292) + N * 200)
    for pc, name in ((1, "jacobi"), (2, "block6")):
        uu, info = ctx.pcg(rowptr, colidx, vbc, b, tol=1e-8, maxiter=20000, precond=pc, profile_iters=64)
        it_bytes = nnzb * 292 + N * 100 + 6 * N * (96 + (48 if pc == 2 else 0))
        print(f"{'PCG (CG form) ' + name:34s} {cfg:28s} iters={info['iters']} info={info['info']} solve={info['solve_ms']:.2f} ms  {6*N*info['iters']/info['solve_ms']/1e6:8.2f} G DOF-it/s  "
              f"spmv={1e3*info['spmv_ms']:.1f} us ({(nnzb*292+N*148)/info['spmv_ms']/1e6:.0f} GB/s frac={(nnzb*292+N*148)/info['spmv_ms']/1e6/peak:.2f}) update={1e3*info['update_ms']:.1f} us  "
              f"iteration(SURVEY bytes)={it_bytes*info['iters']/info['solve_ms']/1e6:.0f} GB/s frac={it_bytes*info['iters']/info['solve_ms']/1e6/peak:.2f}", flush=True)
    # matrix-free operator on the same system: product alone, then the PCG
    ctx.matfree_setup(x, y, z, en0, en1, rad, N, E, NU, fixed=fd)
    ms = timeit(lambda: ctx.matfree_apply(u, out=yv, eliminated=True))
    n_inc = 2 * Ecount
    print(f"{'k_mf_apply (matrix-free y = A u)':34s} {cfg:28s} {6*N/ms/1e6:10.2f} G DOF/s  {ms*1e3:10.1f} us  {n_inc/ms/1e6:6.1f} G incidences/s  "
          f"streamed {(n_inc*32 + N*(32+96))/ms/1e6:.0f} GB/s  [assembled-equivalent {(nnzb*292+N*100)/ms/1e6:.0f} GB/s = {(nnzb*292+N*100)/ms/1e6/peak:.2f} of HBM peak]", flush=True)
    bm = ctx.matfree_rhs(gd, fv)
    for pc, name in ((1, "jacobi"), (2, "block6")):
        uu, info = ctx.pcg_matfree(bm, tol=1e-8, maxiter=20000, precond=pc, profile_iters=64)
        print(f"{'PCG matrix-free ' + name:34s} {cfg:28s} iters={info['iters']} info={info['info']} solve={info['solve_ms']:.2f} ms  {6*N*info['iters']/info['solve_ms']/1e6:8.2f} G DOF-it/s  "
              f"product={1e3*info['spmv_ms']:.1f} us update={1e3*info['update_ms']:.1f} us ({6*N*(88+(28 if pc==2 else 8))/info['update_ms']/1e6:.0f} GB/s frac={6*N*(88+(28 if pc==2 else 8))/info['update_ms']/1e6/peak:.2f})", flush=True)
    del vals, vbc, u, yv
    torch.cuda.empty_cache()

# config 4: BCC 60^3 cells, one Schur complement per cell (n_I = 6 at m=1, 54 at m=2)
rng = np.random.default_rng(44)
radii = 0.02 + 0.06 * rng.random(216000)
for m_ in (1, 2, 3):
    batch, bnd = synthetic_cell_batch(ctx, "BCC", radii, m_, E, NU)
    ms = timeit(lambda: batch.schur(), n=3, warm=1)
    nI = 6 * (batch.xyz.shape[1] - 8)
    print(f"{'k_schur_dense':34s} {'BCC 216000 cells m=%d nI=%d' % (m_, nI):28s} {216000/ms/1e3:10.3f} M cells/s  {ms:8.2f} ms   out {216000*48*48*8/ms/1e6:.0f} GB/s", flush=True)
    if m_ == 1:
        S = batch.schur()
        gidx = torch.from_numpy(rng.integers(0, 2_000_000, size=(216000, 48)).astype(np.int32)).to(dev)
        xx = torch.randn(2_000_000, dtype=torch.float64, device=dev); yy = torch.empty_like(xx)
        ms = timeit(lambda: ctx.ddm_matvec(S, gidx, xx, out=yy))
        line("k_ddm_matvec (S per cell)", "BCC 216000 cells nb=48", 216000, "cells", ms, 216000 * (48 * 48 * 8 + 12 * 48))
        del S
    del batch
    torch.cuda.empty_cache()
