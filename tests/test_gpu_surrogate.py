"""GPU parity of the reduced-basis / surrogate Schur pipeline (csrc/lattice_surrogate.cu) through
the C ABI, against the reference's stored reduced bases, frozen outputs of the reference's own code
(tests/golden/surrogate_ref.npz, made by tests/golden/make_golden_surrogate.py) and the CPU oracle."""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _dataset(name):
    d = np.load(os.path.join(G, f"schur_{name}.npz"))
    return {tuple(r): S for r, S in zip(d["radius_values"], d["schur_matrices"])}


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(G, "surrogate_ref.npz"))


@pytest.fixture(scope="module")
def rb6():
    return np.load(os.path.join(G, "reduced_basis_BCC_tol_1e-6.npz"))


@pytest.mark.parametrize("name,tol,tag", [("BCC", 1e-3, "BCC_tol_1e-3"), ("BCC", 1e-6, "BCC_tol_1e-6"),
                                          ("Hybrid1", 1e-6, "Hybrid1_tol_1e-6"), ("Hybrid4", 1e-6, "Hybrid4_tol_1e-6")])
def test_greedy_reproduces_the_reference_stored_bases(ctx, name, tol, tag):
    from pylatticedso_b200 import surrogate
    rb = np.load(os.path.join(G, f"reduced_basis_{tag}.npz"))
    out = surrogate.reduce_basis_greedy(_dataset(name), tol, verbose=0, ctx=ctx)
    B, A = out[3], out[4]
    assert B.shape == rb["basis_reduced_ortho"].shape
    assert np.abs(B - rb["basis_reduced_ortho"]).max() < 1e-9
    assert np.abs(A - rb["alpha_ortho"]).max() < 1e-9 * np.abs(rb["alpha_ortho"]).max()


def test_greedy_bookkeeping_and_npz_schema(ctx, ref, tmp_path):
    from pylatticedso_b200 import surrogate
    for tag, tol in (("g3", 1e-3), ("g6", 1e-6)):
        main_e, coef, pp, B, A, matP, nrm = surrogate.reduce_basis_greedy(_dataset("BCC"), tol, file_name=f"rb_{tag}", verbose=0,
                                                                        ctx=ctx, save_dir=tmp_path)
        assert (main_e == ref[f"{tag}_mainelem"]).all()
        np.testing.assert_allclose(coef, ref[f"{tag}_reducedcoef"], rtol=0, atol=1e-7 * np.abs(ref[f"{tag}_reducedcoef"]).max())
        np.testing.assert_allclose(matP, ref[f"{tag}_matP"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(nrm, ref[f"{tag}_norms"], rtol=1e-13)
        assert len(pp) == len(main_e) and pp[0].shape == (48, 48)
        back = surrogate.load_reduced_basis(tmp_path / f"rb_{tag}")
        assert set(back.files) == {"basis_reduced_ortho", "alpha_ortho", "list_elements"}
        assert back["list_elements"].shape == (10, 1)
        np.testing.assert_array_equal(back["basis_reduced_ortho"], B)


def test_projection(ctx, ref):
    from pylatticedso_b200 import surrogate
    sd = _dataset("BCC")
    keys = list(sd)[:3]
    al = surrogate.project_to_reduced_basis({k: sd[k] for k in keys}, ref["g6_basis"], ctx=ctx)
    np.testing.assert_allclose(np.stack([al[k] for k in keys]), ref["proj_alphas"], rtol=0, atol=1e-9 * np.abs(ref["proj_alphas"]).max())


def test_rbf_1d_fit_value_gradient(ctx, ref, rb6):
    from pylatticedso_b200 import surrogate
    r = surrogate.ThinPlateSplineRBF(rb6["list_elements"], rb6["alpha_ortho"].T, ctx=ctx)
    # the TPS system is ill conditioned: the weights agree to ~cond * eps, the interpolant far better
    np.testing.assert_allclose(r.W, ref["s1_rbf_W"], rtol=0, atol=1e-6 * np.abs(ref["s1_rbf_W"]).max())
    scale = np.abs(ref["s1_rbf_alphas"]).max()
    np.testing.assert_allclose(r.evaluate(ref["q1"]), ref["s1_rbf_alphas"], rtol=0, atol=1e-10 * scale)
    assert r.evaluate(ref["q1"][0]).shape == (5,) and r.gradient(ref["q1"][0]).shape == (1, 5)


@pytest.mark.parametrize("kind", ["RBF", "linear", "nearest_neighbor"])
def test_schur_batch_equals_the_reference_surrogates(ctx, ref, rb6, kind):
    from pylatticedso_b200 import surrogate
    s = surrogate.SchurSurrogate(rb6, kind, ctx=ctx)
    S = s.schur_batch(ref["q1"])
    want = ref[f"s1_{kind}"]
    assert S.shape == want.shape == (16, 48, 48)
    tol = 1e-10 if kind == "RBF" else 1e-12
    np.testing.assert_allclose(S, want, rtol=0, atol=tol * np.abs(want).max())
    if kind == "RBF":
        dS = s.schur_gradients_device(ref["q1"][:5]).cpu().numpy()
        np.testing.assert_allclose(dS, ref["s1_rbf_dS"], rtol=0, atol=1e-9 * np.abs(ref["s1_rbf_dS"]).max())
    else:
        with pytest.raises(NotImplementedError):
            s.schur_gradients_device(ref["q1"][:1])


def test_rbf_2d_against_the_reference_interpolant(ctx, ref):
    from pylatticedso_b200 import surrogate
    r = surrogate.ThinPlateSplineRBF(ref["x2"], ref["a2"], ctx=ctx)
    np.testing.assert_allclose(r.evaluate(ref["q2"]), ref["r2_eval"], rtol=0, atol=1e-9 * np.abs(ref["r2_eval"]).max())
    np.testing.assert_allclose(r.gradient(ref["q2"]), ref["r2_grad"], rtol=0, atol=1e-9 * np.abs(ref["r2_grad"]).max())


@pytest.mark.parametrize("n,k,M", [(36, 1, 1), (42, 3, 31), (48, 5, 257), (36, 38, 1000), (48, 64, 100), (42, 70, 65), (78, 17, 33)])
def test_basis_expand_dmma_against_numpy(ctx, n, k, M):
    """lat_basis_expand (DMMA) == basis @ alphas with the reference's order='F' reshape; n = 42, 78 exercise the
    column tail (n*n % 16 == 4), k = 70 the accumulate pass, M the row tail."""
    from oracle import surrogate_oracle as so
    from pylatticedso_b200 import surrogate
    rng = np.random.default_rng(n * 1000 + k)
    basis = rng.standard_normal((n * n, k))
    alphas = rng.standard_normal((M, k))
    s = surrogate.SchurSurrogate({"basis_reduced_ortho": basis, "alpha_ortho": np.zeros((k, 2)), "list_elements": np.zeros((2, 1))},
                                 "nearest_neighbor", ctx=ctx)
    got = s.expand_device(surrogate._dev(ctx, alphas)).cpu().numpy()
    want = so.schur_from_alphas(basis, alphas, n)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13 * np.abs(want).max() * k)


def test_config4_size_batch_properties(ctx, rb6):
    """BASELINE config 4 size: 216 000 cells from the BCC basis.  Size-independent checks: a sample against the oracle,
    unit coefficient vectors return the basis vectors, linearity in alpha."""
    import torch
    from oracle import surrogate_oracle as so
    from pylatticedso_b200 import surrogate
    s = surrogate.SchurSurrogate(rb6, "RBF", ctx=ctx)
    M = 216000
    rng = np.random.default_rng(3)
    radii = rng.uniform(0.01, 0.1, (M, 1))
    S = s.schur_batch_device(radii)
    assert S.shape == (M, 48, 48)
    pick = rng.choice(M, 64, replace=False)
    wcp = so.tps_fit(rb6["list_elements"], rb6["alpha_ortho"].T)
    want = so.schur_from_alphas(rb6["basis_reduced_ortho"], so.tps_evaluate(rb6["list_elements"], wcp, radii[pick]), 48)
    got = S[torch.as_tensor(pick, device=S.device)].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-10 * np.abs(want).max())
    assert float((S - S.transpose(1, 2)).abs().max()) < 1e-9 * float(S.abs().max())
    del S
    eye = torch.eye(5, dtype=torch.float64, device=ctx.device)
    E = s.expand_device(eye).cpu().numpy()
    for j in range(5):
        np.testing.assert_array_equal(E[j], rb6["basis_reduced_ortho"][:, j].reshape(48, 48, order="F"))
    a = torch.as_tensor(rng.standard_normal((1000, 5)), device=ctx.device)
    b = torch.as_tensor(rng.standard_normal((1000, 5)), device=ctx.device)
    lin = s.expand_device(a + 2.0 * b) - (s.expand_device(a) + 2.0 * s.expand_device(b))
    assert float(lin.abs().max()) < 1e-13 * 10


def test_latticesim_drop_ins_on_a_duck_typed_lattice(ctx, ref, rb6):
    """The functions install.patch_reference binds to LatticeSim, on an object with the attributes the reference's
    methods read (lattice_sim.py:921-978, 1056-1082)."""
    from pylatticedso_b200 import surrogate
    me = types.SimpleNamespace(type_schur_complement_computation="RBF", reduce_basis_dict=rb6,
                               alpha_coefficients_greedy=rb6["alpha_ortho"].T, radial_basis_function=None,
                               shape_schur_complement=None)
    S = surrogate.lattice_schur_batch(me, [list(x) for x in ref["q1"]], ctx=ctx)
    np.testing.assert_allclose(S, ref["s1_RBF"], rtol=0, atol=1e-10 * np.abs(ref["s1_RBF"]).max())
    assert me.shape_schur_complement == 48 and me.radial_basis_function is not None
    one = surrogate.lattice_schur_single(me, list(ref["q1"][3]), ctx=ctx)
    np.testing.assert_allclose(one, ref["s1_RBF"][3], rtol=0, atol=1e-10 * np.abs(ref["s1_RBF"]).max())
    dS = surrogate.lattice_schur_gradients_rbf(me, list(ref["q1"][2]), ctx=ctx)
    assert isinstance(dS, list) and len(dS) == 1
    np.testing.assert_allclose(dS[0], ref["s1_rbf_dS"][2, 0], rtol=0, atol=1e-9 * np.abs(ref["s1_rbf_dS"]).max())
    surrogate.lattice_define_rbf(me, ctx=ctx)
    np.testing.assert_allclose(me.radial_basis_function.evaluate(ref["q1"]), ref["s1_rbf_alphas"], rtol=0,
                               atol=1e-10 * np.abs(ref["s1_rbf_alphas"]).max())


def test_error_paths_and_edges(ctx, rb6):
    """Loud failures instead of silent garbage: duplicate RBF centres, a rank-deficient basis in the projection, the
    1-D look-up called with two parameters; empty query batches are a no-op."""
    import torch
    from pylatticedso_b200 import surrogate
    from pylatticedso_b200.lib import LatticeB200Error
    with pytest.raises(LatticeB200Error, match="singular"):
        surrogate.ThinPlateSplineRBF(np.array([[0.1], [0.1], [0.3]]), np.array([1.0, 1.0, 2.0]), ctx=ctx)
    B = np.random.default_rng(0).standard_normal((36, 2))
    with pytest.raises(LatticeB200Error, match="rank deficient"):
        surrogate.project_to_reduced_basis({0: np.ones((6, 6))}, np.column_stack([B[:, 0], B[:, 0]]), ctx=ctx)
    with pytest.raises(LatticeB200Error, match="one parameter"):       # the C entry point of the 1-D look-up says so
        x2 = surrogate._dev(ctx, np.zeros((4, 2)))
        ctx.check(ctx.lib.lat_alpha_lookup(ctx.h, 1, surrogate._ptr(x2), 4, 2, surrogate._ptr(x2), 2, surrogate._ptr(x2), 4,
                                           surrogate._ptr(x2)))
    with pytest.raises(NotImplementedError):
        surrogate.SchurSurrogate(rb6, "kriging", ctx=ctx)
    s = surrogate.SchurSurrogate(rb6, "RBF", ctx=ctx)
    empty = s.expand_device(torch.empty((0, 5), dtype=torch.float64, device=ctx.device))
    assert tuple(empty.shape) == (0, 48, 48)
    with pytest.raises(ValueError):
        s.schur_batch([[0.05, 0.06]])                     # two parameters for a one-parameter surrogate


def test_rbf_many_centres_and_three_parameters(ctx):
    """More centres than one shared-memory chunk (128) and d = 3: device fit / value / gradient against the oracle."""
    from oracle import surrogate_oracle as so
    from pylatticedso_b200 import surrogate
    rng = np.random.default_rng(21)
    X = rng.uniform(0.0, 1.0, (300, 3))
    Y = np.stack([np.sin(X @ np.array([1.0, 2.0, 3.0])), X[:, 0] * X[:, 1], np.exp(-X[:, 2])], axis=1)
    r = surrogate.ThinPlateSplineRBF(X, Y, reg=1e-10, ctx=ctx)
    wcp = so.tps_fit(X, Y, reg=1e-10)
    q = rng.uniform(0.0, 1.0, (77, 3))
    np.testing.assert_allclose(r.evaluate(q), so.tps_evaluate(X, wcp, q), rtol=0, atol=1e-7)
    np.testing.assert_allclose(r.gradient(q), so.tps_gradient(X, wcp, q), rtol=0, atol=1e-6)
    np.testing.assert_allclose(r.evaluate(X[:10]), Y[:10], rtol=0, atol=1e-6)        # interpolation at the centres


def test_linear_surrogate_in_two_parameters_against_the_reference(ctx, ref):
    """evaluate_alphas_linear_surrogate, N-parameter branch (lattice_sim.py:794-807): Delaunay interpolation inside the
    hull of the 100 centres of the reference's BCC+Hybrid4 set, nearest centre outside (three of the queries)."""
    from pylatticedso_b200 import surrogate
    k = ref["a2"].shape[1]
    s = surrogate.SchurSurrogate({"basis_reduced_ortho": np.eye(36, k), "alpha_ortho": ref["a2"].T, "list_elements": ref["x2"]},
                                 "linear", ctx=ctx)
    got = s.alphas_device(s._queries(ref["q2l"])).cpu().numpy()
    np.testing.assert_allclose(got, ref["l2_eval"], rtol=0, atol=1e-12 * np.abs(ref["l2_eval"]).max())


def test_linear_surrogate_three_parameters_against_scipy(ctx):
    """k_alpha_simplex in three parameters (tetrahedra) on scattered centres: scipy's LinearNDInterpolator inside the
    hull, nearest centre outside -- the oracle's restatement of lattice_sim.py:794-807."""
    from oracle import surrogate_oracle as so
    from pylatticedso_b200 import surrogate
    rng = np.random.default_rng(17)
    X = rng.uniform(0.0, 1.0, (60, 3))
    A = rng.standard_normal((60, 7))
    s = surrogate.SchurSurrogate({"basis_reduced_ortho": np.eye(36, 7), "alpha_ortho": A.T, "list_elements": X}, "linear", ctx=ctx)
    Q = np.vstack([rng.uniform(0.1, 0.9, (200, 3)), rng.uniform(-0.5, 1.5, (50, 3))])
    got = s.alphas_device(s._queries(Q)).cpu().numpy()
    want = so.alphas_linear_nd(X, A, Q)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-11 * np.abs(want).max())


def test_greedy_on_a_larger_synthetic_snapshot_set_equals_the_oracle(ctx):
    """120 snapshots of a smooth 2-parameter family of symmetric 24x24 matrices: same selected snapshots, basis size,
    basis and coefficients as the oracle's restatement of the reference loop."""
    from oracle import surrogate_oracle as so
    from pylatticedso_b200 import surrogate
    rng = np.random.default_rng(2)
    B0, B1, B2, B3 = (0.5 * (M_ + M_.T) for M_ in rng.standard_normal((4, 24, 24)))
    sd = {}
    for a in np.linspace(0.1, 1.0, 12):
        for b in np.linspace(0.2, 0.8, 10):
            sd[(round(float(a), 6), round(float(b), 6))] = B0 + a * B1 + b * b * B2 + np.sin(3 * a * b) * B3 + 1e-3 * np.exp(a) * np.eye(24)
    got = surrogate.reduce_basis_greedy(sd, 1e-8, verbose=0, ctx=ctx)
    want = so.greedy_reduced_basis(sd, 1e-8)
    assert got[3].shape == want[3].shape and (got[0] == want[0]).all()
    assert np.abs(got[3] - want[3]).max() < 1e-7                       # later basis vectors amplify rounding (deflation)
    # what matters downstream: the reconstruction of every snapshot from (basis, alpha)
    keys = sorted(sd)
    rec = got[3] @ got[4]
    for j, k_ in enumerate(keys):
        S = rec[:, j].reshape(24, 24, order="F")
        assert np.abs(S - sd[k_]).max() < 1e-6 * np.abs(sd[k_]).max()
