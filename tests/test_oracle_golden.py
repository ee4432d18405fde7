"""CPU: the oracle against every known-answer fixture."""
import numpy as np
import pytest

from conftest import E_MOD, NU, load_golden, mesh_from_npz
from oracle import lattice_oracle as orc


@pytest.mark.parametrize("geom,tol", [("BCC", 1e-12), ("Hybrid1", 5e-12), ("Hybrid4", 1e-12)])
def test_oracle_reproduces_reference_schur_goldens(geom, tol):
    """30 Schur matrices stored by the reference (real dolfinx/PETSc runs)."""
    G = load_golden(f"schur_{geom}.npz")
    SM = G["schur_matrices"]
    assert SM.shape[0] == 10
    for i in range(SM.shape[0]):
        m = mesh_from_npz(G, f"c{i}_")
        K = orc.assemble_csr(m.xyz, np.stack([m.en0, m.en1], 1), m.rad, E_MOD, NU)
        S = orc.schur_complement(K, G[f"c{i}_bnd"])
        err = np.abs(S - SM[i]).max() / np.abs(SM[i]).max()
        assert err < tol, (geom, i, err)


def test_element_matrix_properties():
    rng = np.random.default_rng(0)
    x1 = rng.standard_normal((50, 3))
    x2 = x1 + rng.standard_normal((50, 3))
    # include axis-aligned and tie cases of the frame rule
    x1[:4] = 0.0
    x2[0] = [1, 0, 0]; x2[1] = [0, 1, 0]; x2[2] = [0, 0, 1]; x2[3] = [1, 1, 0]
    r = rng.uniform(0.01, 0.1, 50)
    K = orc.element_stiffness(x1, x2, r, E_MOD, NU)
    assert np.abs(K - K.transpose(0, 2, 1)).max() < 1e-9 * np.abs(K).max()
    for k in range(50):
        w = np.linalg.eigvalsh(K[k])
        assert (w > -1e-9 * w.max()).all()
        assert (np.abs(w) > 1e-9 * w.max()).sum() == 6      # rank 6: six rigid-body modes
    # derivative against central differences
    h = 1e-6
    dK = orc.element_stiffness(x1, x2, r, E_MOD, NU, drad=True)
    fd = (orc.element_stiffness(x1, x2, r + h, E_MOD, NU) - orc.element_stiffness(x1, x2, r - h, E_MOD, NU)) / (2 * h)
    assert np.abs(dK - fd).max() < 1e-6 * np.abs(dK).max()


@pytest.mark.parametrize("case", ["disp", "force"])
def test_oracle_fem_matches_reference_ddm_in_the_loop(case):
    """Reference's own solve_DDM (fed with oracle Schur matrices) vs the oracle's full FEM."""
    G = load_golden(f"ddm_loop_{case}.npz")
    m = mesh_from_npz(G)
    K = orc.assemble_csr(m.xyz, np.stack([m.en0, m.en1], 1), m.rad, E_MOD, NU)
    u, R = orc.solve_static(K, G["fixed"].astype(bool), G["g"], G["f"])
    up = u.reshape(-1, 6)[: m.n_points]
    ref = G["u_points_reference_ddm"]
    sel = G["point_on_cell_boundary"]
    assert np.abs(up[sel] - ref[sel]).max() / np.abs(ref).max() < 1e-9
    # global equilibrium of the reactions
    free = ~G["fixed"].astype(bool)
    assert np.abs((R - G["f"])[free]).max() < 1e-8 * np.abs(R).max()


def test_oracle_gradient_matches_reference_optimiser_in_the_loop():
    G = load_golden("grad_loop_bcc311.npz")
    m = mesh_from_npz(G)
    en = np.stack([m.en0, m.en1], 1)
    K = orc.assemble_csr(m.xyz, en, m.rad, E_MOD, NU)
    u, _ = orc.solve_static(K, G["fixed"].astype(bool), G["g"], G["f"])
    assert abs(float(G["f"] @ u) - float(G["compliance_reference"])) < 1e-8 * abs(float(G["compliance_reference"]))
    g = orc.compliance_gradient(m.xyz, en, m.rad, u, G["group"], 3, E_MOD, NU, chain=m.chain)
    ref = G["gradient_reference_fd"]
    # the reference gradient is a central FD of Schur matrices (lattice_sim.py:1020-1054, h = 1e-6):
    # its own noise is ~2e-6 |g|_inf, so 1e-6-level agreement is only meaningful relative to |g|_inf
    assert np.abs(g - ref).max() < 5e-6 * np.abs(ref).max()
    assert np.abs(g - G["gradient_oracle_analytic"]).max() < 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("name", ["well_default", "well_ddm", "ill_clamp", "ill_restart"])
def test_oracle_pcg_is_the_reference_pcg(name):
    G = load_golden("pcg_reference.npz")
    A = G["A_well"] if name.startswith("well") else G["A_ill"]
    maxiter, tol, mintol, restart, amax = G[f"{name}_params"]
    Minv = np.diag(1.0 / np.diag(A)) if bool(G[f"{name}_jacobi"]) else None
    x, info, it = orc.reference_pcg(A, G[f"{name}_b"], Minv, maxiter=int(maxiter), tol=tol, mintol=mintol,
                                    restart_every=int(restart), alpha_max=amax)
    assert info == int(G[f"{name}_info"]) and it == int(G[f"{name}_iters"])
    assert np.array_equal(x, G[f"{name}_x"])


def test_matrix_free_closed_form_equals_element_stiffness():
    """The closed form evaluated by the CUDA matrix-free operator (csrc/matfree.cuh) reproduces K_e u for both
    ends of the element, including axis-aligned struts and the tie cases of the frame rule."""
    rng = np.random.default_rng(3)
    n = 64
    x1 = rng.standard_normal((n, 3))
    x2 = x1 + rng.standard_normal((n, 3))
    x1[:4] = 0.0
    x2[0] = [1, 0, 0]; x2[1] = [0, 1, 0]; x2[2] = [0, 0, -2]; x2[3] = [1, 1, 0]
    r = rng.uniform(0.01, 0.1, n)
    u = rng.standard_normal((n, 12))
    K = orc.element_stiffness(x1, x2, r, E_MOD, NU)
    f = np.einsum("nij,nj->ni", K, u)
    f0 = orc.element_action_closed_form(x1, x2, r, u[:, :6], u[:, 6:], E_MOD, NU)
    f1 = orc.element_action_closed_form(x2, x1, r, u[:, 6:], u[:, :6], E_MOD, NU)
    scale = np.abs(f).max(axis=1, keepdims=True)
    assert (np.abs(f0 - f[:, :6]) <= 1e-12 * scale).all()
    assert (np.abs(f1 - f[:, 6:]) <= 1e-12 * scale).all()


@pytest.mark.parametrize("geom,tol", [("BCC", 1e-12), ("Hybrid1", 5e-12), ("Hybrid4", 1e-12)])
def test_chain_condensation_reproduces_reference_schur_goldens(geom, tol):
    """The strut pre-pass restated on the CPU (spring series + planar-beam chain condensation, joint-only cell)
    against the reference's 30 stored Schur matrices -- the checker of k_chain_condense."""
    G = load_golden(f"schur_{geom}.npz")
    SM = G["schur_matrices"]
    for i in range(SM.shape[0]):
        m = mesh_from_npz(G, f"c{i}_")
        bnd_nodes = G[f"c{i}_bnd"].reshape(-1, 6)[:, 0] // 6
        S, chains = orc.schur_via_chain_condensation(m.xyz, np.stack([m.en0, m.en1], 1), m.rad, bnd_nodes, E_MOD, NU)
        assert max(len(c[2]) for c in chains) > 1            # the goldens do have multi-element struts
        err = np.abs(S - SM[i]).max() / np.abs(SM[i]).max()
        assert err < tol, (geom, i, err)


def test_condensed_strut_equals_dense_condensation():
    """Non-uniform element lengths and radii, flipped elements: condensed 12x12 == dense Schur of the chain."""
    rng = np.random.default_rng(5)
    m = 9
    A = np.array([0.1, 0.2, 0.3]); B = A + np.array([0.5, -0.4, 0.7])
    fr = np.sort(rng.random(m - 1))
    pts = np.vstack([A] + [A + f * (B - A) for f in fr] + [B])
    en = np.array([[k, k + 1] for k in range(m)])
    en[2] = en[2][::-1]; en[6] = en[6][::-1]
    rad = rng.uniform(0.02, 0.08, m)
    chain = (0, m, [(k, bool(en[k][0] != k)) for k in range(m)])
    Ks, _ = orc.condensed_strut(pts, en, rad, chain, E_MOD, NU)
    K = orc.assemble_csr(pts, en, rad, E_MOD, NU).toarray()
    bd = np.r_[0:6, 6 * m:6 * m + 6]
    it = np.setdiff1d(np.arange(6 * (m + 1)), bd)
    Sd = K[np.ix_(bd, bd)] - K[np.ix_(bd, it)] @ np.linalg.solve(K[np.ix_(it, it)], K[np.ix_(it, bd)])
    assert np.abs(Ks - Sd).max() <= 1e-12 * np.abs(Sd).max()


@pytest.mark.parametrize("geom,cells,mseg", [("BCC", (2, 2, 2), 5), ("Octet", (2, 1, 2), 3)])
def test_joint_only_system_equals_full_system_at_the_joints(geom, cells, mseg):
    """Static condensation of every strut (checker of lat_assemble_bsr_struts): with loads and constraints on
    lattice points only, the joint-only solve has the joint displacements and reactions of the full solve."""
    from pylatticedso_b200 import mesh as M
    lat = M.synthetic_lattice(geom, cells, [0.04], grad_radius=("linear", [True, False, True], [0.01, 0, 0.005]))
    mesh = M.mesh_from_synthetic(lat, mseg)
    fixed, g, f = M.compression_bc(mesh)
    f = f.copy(); f[6 * 3 + 1] = 0.02
    en = np.stack([mesh.en0, mesh.en1], 1)
    K = orc.assemble_csr(mesh.xyz, en, mesh.rad, E_MOD, NU)
    u, R = orc.solve_static(K, fixed.astype(bool), g, f)
    nj = 6 * mesh.n_points
    assert not fixed[nj:].any() and not f[nj:].any()
    Kj, chains = orc.assemble_joint_only(mesh.xyz, en, mesh.rad, mesh.n_points, E_MOD, NU)
    assert len(chains) * mseg == mesh.n_elems
    uj, Rj = orc.solve_static(Kj, fixed[:nj].astype(bool), g[:nj], f[:nj])
    assert np.abs(u[:nj] - uj).max() <= 1e-10 * np.abs(u).max()
    assert np.abs(R[:nj] - Rj).max() <= 1e-10 * np.abs(R).max()



def test_ill_restart_iterate_is_rounding_sensitive_at_the_1e4_level():
    """Why the GPU parity test accepts 1e-3 on x for `ill_restart` only (tests/test_gpu_parity.py): the case runs 300
    NON-converged iterations on a matrix with cond ~ 2e8, and on the CPU itself a rounding-level change -- b perturbed
    by 1e-16 relative, or the SAME matrix-vector product summed in a different column order -- moves the final
    iterate by 1e-5..1e-4 relative while iteration count and info code stay put.  Bit-level agreement of x is not a
    property of the algorithm here; iteration count, info code and a 1e-3 band are what can be compared."""
    G = load_golden("pcg_reference.npz")
    A, b = G["A_ill"], G["ill_restart_b"]
    maxiter, tol, mintol, restart, amax = G["ill_restart_params"]
    d = 1.0 / np.diag(A)
    Minv = (lambda v: d * v) if bool(G["ill_restart_jacobi"]) else None
    kw = dict(maxiter=int(maxiter), tol=tol, mintol=mintol, restart_every=int(restart), alpha_max=amax)
    x0, info0, it0 = orc.reference_pcg(A, b, Minv, **kw)
    assert np.linalg.cond(A) > 1e8 and it0 == 300 and info0 == int(G["ill_restart_info"])
    rng = np.random.default_rng(0)
    perm = rng.permutation(A.shape[0])
    Ac = np.ascontiguousarray(A[:, perm])
    x1, info1, it1 = orc.reference_pcg(lambda v: Ac @ v[perm], b, Minv, **kw)           # same product, other summation order
    x2, info2, it2 = orc.reference_pcg(A, b * (1.0 + 1e-16 * rng.standard_normal(b.shape)), Minv, **kw)
    assert (info1, it1) == (info0, it0) and (info2, it2) == (info0, it0)
    dev = max(np.abs(x1 - x0).max(), np.abs(x2 - x0).max()) / np.abs(x0).max()
    assert 1e-7 < dev < 1e-3
