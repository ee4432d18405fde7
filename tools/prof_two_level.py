"""One short two-level matrix-free solve for an ncu launch list:  python tools/prof_two_level.py [n] [n_agg] [maxiter]"""
import sys

sys.path.insert(0, ".")
from pylatticedso_b200 import lib as L
from pylatticedso_b200 import mesh as M
from pylatticedso_b200.fem import BeamFEM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n_agg = int(sys.argv[2]) if len(sys.argv) > 2 else 512
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 12
ctx = L.Context()
lat = M.synthetic_lattice("Octet", (n, n, n), [0.03])
m = M.mesh_from_synthetic(lat, 1)
fixed, g, f = M.compression_bc(m)
fem = BeamFEM(m, 1013.0, 0.3, ctx=ctx)
fem.build_pattern()
tl = fem.two_level(fixed, n_agg)
u, _, info = fem.solve_matrix_free(fixed, g, f, tol=1e-8, maxiter=maxiter, want_reactions=False, two_level=tl, check_every=4)
print(n, tl.n_agg, info)
