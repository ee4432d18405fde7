"""Lattice numbering on the host (numpy) against the same on the GPU (torch ops): bit-identity and time.
    python tools/ab_generator_device.py [n]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pylatticedso_b200 import distributed as D
from pylatticedso_b200 import mesh as M

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
for geom, nn, kw in (("Octet", (24, 10, 10), {"i_range": (5, 13)}), ("BCC", (12, 12, 12), {})):
    a = M.synthetic_lattice(geom, nn, [0.03], **kw)
    b = M.synthetic_lattice(geom, nn, [0.03], device=dev, **kw)
    print(geom, nn, "bit-identical:", all(np.array_equal(getattr(a, k), getattr(b, k)) for k in ("pxyz", "b_p1", "b_p2", "b_rad", "b_cell", "b_type")))
for rep in range(2):
    t0 = time.perf_counter(); lm_h, _ = D.generate_slab("Octet", (n, n, n), [0.03], 1, 0, 1); t1 = time.perf_counter()
    lm_d, _ = D.generate_slab("Octet", (n, n, n), [0.03], 1, 0, 1, device=dev); torch.cuda.synchronize(); t2 = time.perf_counter()
    same = all(np.array_equal(getattr(lm_h, k), getattr(lm_d, k)) for k in ("x", "y", "z", "en0", "en1", "rad"))
    print(f"generate_slab Octet {n}^3 ({lm_h.n_nodes} nodes, {lm_h.n_elems} elements): host {t1 - t0:.2f} s, numbering on the GPU {t2 - t1:.2f} s, identical: {same}", flush=True)
