"""K_e generation (row A1) timing: lat_elem_stiffness on large element batches (CUDA events, 3 warm-up + 10 timed)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylatticedso_b200 import lib as L, mesh as M
E, NU = 1013.0, 0.3
ctx = L.Context(); dev = ctx.device
t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(dev)
for geom, n, r in (("BCC", 60, 0.05), ("Octet", 40, 0.03)):
    lat = M.synthetic_lattice(geom, (n, n, n), [r]); mesh = M.mesh_from_synthetic(lat, 1)
    x, y, z, en0, en1, rad = t(mesh.x, np.float64), t(mesh.y, np.float64), t(mesh.z, np.float64), t(mesh.en0, np.int32), t(mesh.en1, np.int32), t(mesh.rad, np.float64)
    Ec = mesh.n_elems
    buf = torch.empty(Ec * 144 + 2, dtype=torch.float64, device=dev)
    res = {}
    for name, off in (("aligned: thread per element, 256-bit stores", 0), ("misaligned: per-entry fallback kernel", 1)):
        Ke = buf[off: off + Ec * 144]
        call = lambda: ctx.check(ctx.lib.lat_elem_stiffness(ctx.h, L._ptr(x), L._ptr(y), L._ptr(z), L._ptr(en0), L._ptr(en1), L._ptr(rad), Ec, E, NU, 0.9, 0, L._ptr(Ke)))
        for _ in range(3): call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = Ke.clone()
        print(f"{geom} {n}^3 E={Ec}  {name:48s} {ms*1e3:8.1f} us  {Ec/ms/1e6:6.2f} G elements/s  {Ec*1216/ms/1e6:6.0f} GB/s (1216 B/element) = {Ec*1216/ms/1e6/6554.6:.2f} of HBM peak", flush=True)
    a, b = list(res.values())
    print("   max |difference| between the two kernels:", float((a - b).abs().max()), flush=True)
