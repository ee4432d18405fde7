"""GPU: the drop-in layer (flatten -> C ABI -> write-back onto the objects) on stand-in lattice objects."""
import numpy as np
import pytest

from conftest import E_MOD, NU
from fake_lattice import FakeLattice
from oracle import lattice_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("matrix_free", [False, True, "condensed"])
def test_solve_fem_dropin_writes_back_like_the_reference(ctx, matrix_free):
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.fem import solve_FEM_B200
    lat = FakeLattice("BCC", (3, 2, 2), 0.05)
    lat.compression()
    if matrix_free == "condensed":
        xsol, model = solve_FEM_B200(lat, elements_per_strut=2, tol=1e-12, ctx=ctx, condense_struts=True)
    else:
        xsol, model = solve_FEM_B200(lat, elements_per_strut=2, tol=1e-12, ctx=ctx, matrix_free=matrix_free)
    assert model.info["info"] in (0, 5)
    mesh = M.mesh_from_synthetic(lat.syn, 2)
    fixed, g, f = M.compression_bc(mesh)
    K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
    uo, Ro = orc.solve_static(K, fixed.astype(bool), g, f)
    uo, Ro = uo.reshape(-1, 6), Ro.reshape(-1, 6)
    for p in lat.points:
        assert np.abs(np.array(p.displacement_vector) - uo[p.index]).max() < 1e-8 * np.abs(uo).max()
        if any(p.fixed_DOF):
            k = sum(1 for c in lat.cells if p in c.points_cell)   # reactions accumulate once per containing cell
            assert np.abs(np.array(p.reaction_force_vector) - k * Ro[p.index]).max() < 1e-8 * np.abs(Ro).max() * k
    xs, idx = lat.get_global_displacement()
    assert np.array_equal(xsol, xs) and len(idx) == len(xs)


def test_get_schur_complement_dropin(ctx):
    from pylatticedso_b200.schur import get_schur_complement
    from pylatticedso_b200.mesh import flatten_lattice, cell_boundary_dofs
    lat = FakeLattice("BCC", (2, 1, 1), 0.04)
    with pytest.raises(ValueError):
        get_schur_complement(lat, None, elements_per_strut=3, ctx=ctx)      # utils_schur.py:35-36
    S = get_schur_complement(lat, 1, elements_per_strut=3, ctx=ctx)
    mesh = flatten_lattice(lat, 1, 3)
    bnd = cell_boundary_dofs(lat.cells[1], mesh)
    K = orc.assemble_csr(mesh.xyz, np.stack([mesh.en0, mesh.en1], 1), mesh.rad, E_MOD, NU)
    So = orc.schur_complement(K, bnd)
    assert S.shape == (48, 48) and np.abs(S - So).max() < 1e-11 * np.abs(So).max()


def test_solve_ddm_dropin_matches_fem(ctx):
    from pylatticedso_b200.ddm import solve_DDM_B200
    from pylatticedso_b200.fem import solve_FEM_B200
    from pylatticedso_b200.schur import get_schur_complement
    a = FakeLattice("BCC", (2, 2, 2), 0.05); a.compression()
    b = FakeLattice("BCC", (2, 2, 2), 0.05); b.compression()
    x_fem, _ = solve_FEM_B200(a, elements_per_strut=2, tol=1e-13, ctx=ctx)
    S = get_schur_complement(b, 0, elements_per_strut=2, ctx=ctx)            # identical cells share one matrix
    for c in b.cells:
        c.schur_complement = S
    x_ddm, info, idx, rhs = solve_DDM_B200(b, tol=1e-13, ctx=ctx)
    assert info in (0, 5) and x_ddm.shape == x_fem.shape
    # compare_FEM_DDM.py:37-38
    assert np.linalg.norm(x_fem - x_ddm) / np.linalg.norm(x_fem) < 1e-8


@pytest.mark.parametrize("name", ["well_default", "ill_clamp"])
def test_conjugate_gradient_solver_dropin(ctx, name):
    """Same call as the reference's conjugate_gradient_solver(A, b, M, maxiter, tol, mintol, restart_every,
    alpha_max) -> (x, info), on a device operator; outputs frozen from the reference's own solver."""
    import torch
    from conftest import load_golden
    from pylatticedso_b200.pcg import BsrOperator, Jacobi, conjugate_gradient_solver
    G = load_golden("pcg_reference.npz")
    A = G["A_well"] if name.startswith("well") else G["A_ill"]
    nb = A.shape[0] // 6
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)
    op = BsrOperator(ctx, t(np.arange(0, nb * nb + 1, nb), np.int32), t(np.tile(np.arange(nb), nb), np.int32),
                     t(A.reshape(nb, 6, nb, 6).transpose(0, 2, 1, 3).reshape(-1), np.float64))
    maxiter, tol, mintol, restart, amax = G[f"{name}_params"]
    seen = []
    x, info = conjugate_gradient_solver(op, G[f"{name}_b"], M=Jacobi if bool(G[f"{name}_jacobi"]) else None,
                                        maxiter=int(maxiter), tol=tol, mintol=mintol, restart_every=int(restart),
                                        alpha_max=amax, callback=seen.append)
    assert info == int(G[f"{name}_info"]) and len(seen) == 1
    assert np.abs(x - G[f"{name}_x"]).max() < 1e-9 * np.abs(G[f"{name}_x"]).max()
    assert np.abs(op @ x - A @ x).max() < 1e-12 * np.abs(A @ x).max()
    with pytest.raises(TypeError):
        conjugate_gradient_solver(A, G[f"{name}_b"])          # a host matrix is not accepted: no CPU fallback


def test_compliance_gradient_lattice_parameter_order(ctx):
    """Parameter order cell.index * n_geom + j (lattice_opti.py:758-761) and the value against the oracle."""
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.fem import compliance_gradient_lattice, solve_FEM_B200
    lat = FakeLattice("BCC", (3, 1, 1), 0.05)
    for p in lat.points:                          # clamp Xmin, tip load on Xmax (dedup: every load once)
        if p.x == 0.0:
            p.fixed_DOF = [True] * 6
        elif p.x == 3.0:
            p.applied_force[2] = -0.025
    xsol, model = solve_FEM_B200(lat, elements_per_strut=3, tol=1e-13, dedup_point_loads=True, ctx=ctx)
    g = compliance_gradient_lattice(lat, model)
    mesh = model.fem.mesh
    en = np.stack([mesh.en0, mesh.en1], 1)
    go = orc.compliance_gradient(mesh.xyz, en, mesh.rad, model.u.cpu().numpy(), lat.syn.b_cell[mesh.beam_of_elem], 3,
                                 E_MOD, NU, chain=mesh.chain)
    assert g.shape == (3,) and np.abs(g - go).max() < 1e-9 * np.abs(go).max()
    assert (g < 0).all() and abs(g[0]) > abs(g[1]) > abs(g[2])     # thicker struts near the clamp help most


@pytest.mark.parametrize("case", ["bcc322_pen", "octet223_graded"])
@pytest.mark.parametrize("mode", ["assembled", "matrix_free", "condensed"])
def test_dropin_on_dumped_reference_objects(ctx, case, mode):
    """solve_FEM_B200 on objects rebuilt from a DUMP of the real pyLatticeDSO object graph (penalised BCC, graded
    Octet; reference gmsh subdivision): Point.displacement_vector, the k-fold accumulated reaction_force_vector and
    xsol must equal what the reference's own write-back methods left on the reference objects (oracle solution)."""
    from conftest import load_golden
    from fake_lattice import lattice_from_dump
    from pylatticedso_b200.fem import solve_FEM_B200
    G = load_golden(f"objgraph_{case}.npz")
    lat = lattice_from_dump(G)
    xsol, model = solve_FEM_B200(lat, elements_per_strut="gmsh", tol=1e-12, ctx=ctx, matrix_free=(mode == "matrix_free"),
                                 condense_struts=(mode == "condensed"))
    assert model.info["info"] in (0, 5)
    u_ref, R_ref = G["u_points_expected"], G["reaction_points_expected"]
    u = np.array([lat.points[int(i)].displacement_vector for i in G["p_index"]])
    R = np.array([lat.points[int(i)].reaction_force_vector for i in G["p_index"]])
    assert np.abs(u - u_ref).max() <= 1e-8 * np.abs(u_ref).max()
    clamped = G["p_fixed"].any(axis=1)
    assert np.abs(R[clamped] - R_ref[clamped]).max() <= 1e-8 * np.abs(R_ref).max()
    assert xsol.shape == G["xsol_expected"].shape
    assert np.abs(xsol - G["xsol_expected"]).max() <= 1e-8 * np.abs(G["xsol_expected"]).max()
    assert np.array_equal(np.asarray(lat.global_displacement_index), G["global_displacement_index"])


def test_cell_quadform_and_parameter_mapping(ctx):
    """lat_cell_quadform + cell_sensitivities_to_parameters against a numpy evaluation of the reference's loops
    (lattice_opti.py:752-839) for unit_cell, constant (hybrid / not) and linear parameterisations."""
    import torch
    from pylatticedso_b200 import ddm
    from pylatticedso_b200.fem import cell_sensitivities_to_parameters

    class P:
        def __init__(self, d): self.d = d

    class Cc:
        def __init__(self, index, center, u, dS):
            self.index, self.center_point, self._u, self.schur_complement_gradient = index, center, u, dS
            self.node_in_order_simulation = [P(u[6 * k: 6 * k + 6]) for k in range(len(u) // 6)]
        def get_displacement_at_nodes(self, nodes): return [n.d for n in nodes]

    class Lat:
        pass
    rng = np.random.default_rng(3)
    nb, ng = 48, 2
    shared = [rng.standard_normal((nb, nb)) for _ in range(3)]
    cells = [Cc(k, (k + 0.5, (k % 2) + 0.5, 0.5), rng.standard_normal(nb), [shared[k % 3], shared[(k + 1) % 3]]) for k in range(7)]
    lat = Lat(); lat.cells = cells
    q = ddm.compliance_gradient_cells(lat, ctx=ctx)
    qo = np.array([[c._u @ (dS @ c._u) for dS in c.schur_complement_gradient] for c in cells])
    assert q.shape == (7, ng) and np.abs(q - qo).max() <= 1e-12 * np.abs(qo).max()
    lam = [rng.standard_normal(nb) for _ in cells]
    qa = ddm.compliance_gradient_cells(lat, ctx=ctx, adjoint=lam)
    qao = np.array([[l @ (dS @ c._u) for dS in c.schur_complement_gradient] for c, l in zip(cells, lam)])
    assert np.abs(qa - qao).max() <= 1e-12 * np.abs(qao).max()
    # parameter mappings
    lat.optimization_parameters = {"type": "unit_cell"}
    g = cell_sensitivities_to_parameters(lat, q)
    assert np.allclose(g, qo.ravel(), rtol=1e-12)
    lat.optimization_parameters = {"type": "constant", "hybrid": True}
    assert np.allclose(cell_sensitivities_to_parameters(lat, q), qo.sum(0), rtol=1e-12)
    lat.optimization_parameters = {"type": "constant", "hybrid": False}; lat.number_parameters = 1
    assert np.allclose(cell_sensitivities_to_parameters(lat, q), [qo.sum()], rtol=1e-12)
    lat.optimization_parameters = {"type": "linear", "direction": ["x", "y"]}; lat.number_parameters = 3
    lat.actual_optimization_parameters = [0.01, -0.005, 0.03]; lat.min_radius, lat.max_radius = 0.02, 0.075
    exp = np.zeros(3)
    for c, row in zip(cells, qo):
        r_un = 0.01 * c.center_point[0] - 0.005 * c.center_point[1] + 0.03
        if 0.02 + 1e-12 < r_un < 0.075 - 1e-12:
            exp += row.sum() * np.array([c.center_point[0], c.center_point[1], 1.0])
    got = cell_sensitivities_to_parameters(lat, q)
    assert np.allclose(got, exp, rtol=1e-12) and np.any(exp != 0)


def _ddm_dump():
    from conftest import load_golden
    from fake_lattice import DumpDdmLattice
    G = load_golden("objgraph_ddm_bcc322.npz")
    return G, DumpDdmLattice(G)


def test_ddm_operator_equals_the_reference_python_loop(ctx):
    """A8 against the reference's OWN calculate_reaction_force_global (frozen output on a random vector), through
    both device forms: the assembled interface matrix (pcg.DdmOperator) and the batched per-cell GEMV (lat_ddm_matvec)."""
    import torch
    from pylatticedso_b200.pcg import DdmOperator
    G, lat = _ddm_dump()
    op = DdmOperator(lat, ctx=ctx)
    y = op @ G["v"]
    assert np.abs(y - G["y_reference"]).max() <= 1e-10 * np.abs(G["y_reference"]).max()
    pos = {int(i): k for k, i in enumerate(G["p_index"])}
    gidx = np.array([[G["p_free_index"][pos[int(n)], d] for n in row for d in range(6)] for row in G["cell_node_order"]], dtype=np.int32)
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a, dtype=d)).to(ctx.device)
    y2 = ctx.ddm_matvec(t(G["schur_shared"], np.float64), t(gidx, np.int32), t(G["v"], np.float64), n_free=int(G["free_DOF"]))
    assert np.abs(y2.cpu().numpy() - G["y_reference"]).max() <= 1e-10 * np.abs(G["y_reference"]).max()
    assert lat.python_loop_calls == 0


def test_conjugate_gradient_solver_accepts_the_reference_linear_operator(ctx):
    """conjugate_gradient_solver(LinearOperator(matvec=lattice.calculate_reaction_force_global), b, M=...) exactly as
    LatticeSim.solve_DDM calls it (lattice_sim.py:1148-1160): the lattice is recognised, the iteration runs on the
    device, and the result equals the reference's own solve_DDM solution."""
    from scipy.sparse.linalg import LinearOperator
    from pylatticedso_b200.pcg import conjugate_gradient_solver
    G, lat = _ddm_dump()
    n = int(G["free_DOF"])
    # (dtype given: without it scipy probes matvec once at construction -- in the reference that is one pass of the
    #  Python loop before conjugate_gradient_solver is even called, nothing the solver can avoid)
    A = LinearOperator(shape=(n, n), matvec=lat.calculate_reaction_force_global, dtype=float)
    seen = []
    x, info = conjugate_gradient_solver(A, G["b_reference"], M=object(), maxiter=2000, tol=1e-11, mintol=1e-14,
                                        restart_every=500000, alpha_max=100, callback=seen.append)
    assert info == 0 and len(seen) == 1 and lat.python_loop_calls == 0
    pos = {int(i): k for k, i in enumerate(G["p_index"])}
    free_of = np.array([G["p_free_index"][pos[int(nn)], d] for nn, d in G["xsol_order"]])
    ref = G["xsol_reference"]
    assert np.abs(x[free_of] - ref).max() <= 1e-7 * np.abs(ref).max()


def test_solve_ddm_b200_equals_the_reference_solve_ddm(ctx):
    from pylatticedso_b200.ddm import solve_DDM_B200
    G, lat = _ddm_dump()
    xsol, info, gdi, b = solve_DDM_B200(lat, tol=1e-12, ctx=ctx)
    assert info == 0 and lat.python_loop_calls == 0
    assert np.abs(xsol - G["xsol_reference"]).max() <= 1e-7 * np.abs(G["xsol_reference"]).max()
    assert np.array_equal(np.asarray(gdi), G["global_displacement_index"])
    # right-hand side: same entries as the reference's b = f_free - r_free (ordered by interface node here)
    assert abs(np.linalg.norm(b) - np.linalg.norm(G["b_reference"])) <= 1e-10 * np.linalg.norm(G["b_reference"])
    ub = np.array([lat.points[int(i)].displacement_vector for i in G["boundary_point_index"]])
    assert np.abs(ub - G["u_boundary_reference"]).max() <= 1e-7 * np.abs(G["u_boundary_reference"]).max()


def test_solve_ddm_b200_two_level_preconditioner(ctx):
    """The same reference solve_DDM dump with the two-level preconditioner standing in for the reference's SuperLU
    preconditioner (lattice_sim.py:1333-1415): the solution does not depend on the preconditioner."""
    from pylatticedso_b200.ddm import solve_DDM_B200
    G, lat = _ddm_dump()
    xsol, info, gdi, b = solve_DDM_B200(lat, tol=1e-12, ctx=ctx, two_level=4)
    assert info == 0 and lat.python_loop_calls == 0
    assert np.abs(xsol - G["xsol_reference"]).max() <= 1e-7 * np.abs(G["xsol_reference"]).max()
    assert np.array_equal(np.asarray(gdi), G["global_displacement_index"])


def test_schur_gradients_dropin_on_dumped_reference_cell(ctx):
    """schur.schur_gradients (the rebinding of LatticeSim._compute_schur_gradients, lattice_sim.py:1020-1054) on a
    penalised BCC cell rebuilt from the reference dump: analytic dS/dr against the reference's recipe -- a central
    finite difference of the cell Schur complement with every beam of the cell set to r (x1.5 on beam_mod segments)."""
    from conftest import load_golden
    from fake_lattice import lattice_from_dump
    from pylatticedso_b200.schur import get_schur_complement, schur_gradients
    G = load_golden("objgraph_bcc322_pen.npz")
    lat = lattice_from_dump(G)
    cell = lat.cells[5]
    r0 = 0.05
    dS = schur_gradients(lat, cell, [r0], elements_per_strut="gmsh", ctx=ctx)
    assert len(dS) == 1 and dS[0].shape == (48, 48)

    def schur_at(r):
        for b in cell.beams_cell:
            b.radius = r * (b.penalization_coefficient if b.beam_mod else 1.0)
        return orc.cell_schur_from_lattice(lat, cell.index, E_MOD, NU, "gmsh")
    h = 1e-6
    fd = (schur_at(r0 + h) - schur_at(r0 - h)) / (2 * h)
    schur_at(r0)
    assert np.abs(dS[0] - fd).max() <= 2e-6 * np.abs(fd).max()
    S = get_schur_complement(lat, cell.index, elements_per_strut="gmsh", ctx=ctx)
    So = orc.cell_schur_from_lattice(lat, cell.index, E_MOD, NU, "gmsh")
    assert np.abs(S - So).max() <= 1e-11 * np.abs(So).max()
