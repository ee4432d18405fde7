"""Block-Jacobi vs two-level (block-Jacobi + rigid-body-mode coarse space) PCG: iterations and time to solution.

    python tools/ab_two_level.py [--big]        # --big adds the 100^3 octet lattice (BASELINE configs[4])
"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pylatticedso_b200 import lib as L
from pylatticedso_b200 import mesh as M
from pylatticedso_b200.fem import BeamFEM

E_MOD, NU = 1013.0, 0.3


def run(ctx, geom, n, m_el, r, targets, matfree, tol=1e-8):
    lat = M.synthetic_lattice(geom, (n, n, n), [r])
    m = M.mesh_from_synthetic(lat, m_el)
    fixed, g, f = M.compression_bc(m)
    fem = BeamFEM(m, E_MOD, NU, ctx=ctx)
    fem.build_pattern()
    solve = fem.solve_matrix_free if matfree else fem.solve
    kw = {} if matfree else {"persistent": False}
    name = f"{geom} {n}^3 m={m_el} ({m.n_dof} DOF, {'matrix-free' if matfree else 'assembled'})"
    u0, _, i0 = solve(fixed, g, f, tol=tol, want_reactions=False, **kw)
    u0, _, i0 = solve(fixed, g, f, tol=tol, want_reactions=False, **kw)
    print(f"{name}: block-Jacobi {i0['iters']} it, {i0['solve_ms']:.1f} ms ({1e3 * i0['solve_ms'] / max(1, i0['iters']):.1f} us/it)", flush=True)
    if not matfree:
        up, _, ip = fem.solve(fixed, g, f, tol=tol, want_reactions=False)
        up, _, ip = fem.solve(fixed, g, f, tol=tol, want_reactions=False)
        if ip["persistent"]:
            print(f"    persistent on-chip kernel: {ip['iters']} it, {ip['solve_ms']:.1f} ms", flush=True)
    for t in targets:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tl = fem.two_level(fixed, t)
        torch.cuda.synchronize()
        t_setup = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        tl2 = fem.two_level(fixed, t)          # warm (allocations, cuSOLVER handles)
        torch.cuda.synchronize()
        t_setup2 = (time.perf_counter() - t0) * 1e3
        u, _, i = solve(fixed, g, f, tol=tol, want_reactions=False, two_level=tl2, **({} if matfree else {}))
        u, _, i = solve(fixed, g, f, tol=tol, want_reactions=False, two_level=tl2)
        du = float((u - u0).abs().max() / u0.abs().max())
        print(f"    two-level n_agg={tl2.n_agg:5d} (n_c={6 * tl2.n_agg}): {i['iters']} it, {i['solve_ms']:.1f} ms "
              f"({1e3 * i['solve_ms'] / max(1, i['iters']):.1f} us/it), set-up {t_setup2:.1f} ms (first {t_setup:.1f}), "
              f"true_relres {i['true_relres']:.2e}, |u-u_bj|/|u| {du:.1e}, speed-up x{i0['solve_ms'] / i['solve_ms']:.2f}", flush=True)
        del tl, tl2
    del fem
    torch.cuda.empty_cache()


def run_condensed(ctx, geom, n, m_el, r, targets, tol=1e-8):
    """The exact joint-only system (struts condensed) with and without the coarse space."""
    lat = M.synthetic_lattice(geom, (n, n, n), [r])
    m = M.mesh_from_synthetic(lat, m_el)
    fixed, g, f = M.compression_bc(m)
    fem = BeamFEM(m, E_MOD, NU, ctx=ctx)
    for _ in range(2):
        u0, _, i0 = fem.solve_condensed(fixed, g, f, tol=tol)
    print(f"{geom} {n}^3 m={m_el} joint-only ({i0['n_dof_condensed']} of {i0['n_dof_full']} DOF): block-Jacobi {i0['iters']} it, "
          f"{i0['solve_ms']:.2f} ms (persistent kernel: {i0['persistent']})", flush=True)
    for t in targets:
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            u, _, i = fem.solve_condensed(fixed, g, f, tol=tol, two_level=t)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
        du = float((u - u0).abs().max() / u0.abs().max())
        print(f"    two-level target {t}: {i['iters']} it, {i['solve_ms']:.2f} ms PCG ({wall:.1f} ms wall for pattern + condensation + "
              f"coarse set-up + solve), |u-u_bj|/|u| {du:.1e}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    a = ap.parse_args()
    ctx = L.Context()
    run(ctx, "BCC", 20, 2, 0.05, [216, 512, 1000], False)
    run(ctx, "Octet", 20, 1, 0.03, [216, 512, 1000], False)
    run(ctx, "Octet", 40, 1, 0.03, [512, 1000, 2048], False)
    run(ctx, "Octet", 40, 1, 0.03, [512, 1000, 2048], True)
    run_condensed(ctx, "BCC", 20, 2, 0.05, [64, 216])
    run_condensed(ctx, "BCC", 40, 4, 0.05, [216, 512])
    if a.big:
        run(ctx, "Octet", 100, 1, 0.03, [512, 1000, 2048], True)


if __name__ == "__main__":
    main()
