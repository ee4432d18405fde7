"""CPU: the surrogate oracle (oracle/surrogate_oracle.py) against the reference's stored reduced bases and against
frozen outputs of the reference's own greedy / RBF / surrogate-evaluation code (tests/golden/surrogate_ref.npz)."""
import os

import numpy as np
import pytest

from oracle import surrogate_oracle as so

G = os.path.join(os.path.dirname(__file__), "golden")


def _dataset(name):
    d = np.load(os.path.join(G, f"schur_{name}.npz"))
    return {tuple(r): S for r, S in zip(d["radius_values"], d["schur_matrices"])}


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(G, "surrogate_ref.npz"))


@pytest.mark.parametrize("name,tol,tag", [("BCC", 1e-3, "BCC_tol_1e-3"), ("BCC", 1e-6, "BCC_tol_1e-6"),
                                          ("Hybrid1", 1e-6, "Hybrid1_tol_1e-6"), ("Hybrid4", 1e-6, "Hybrid4_tol_1e-6")])
def test_greedy_reproduces_the_reference_stored_bases(name, tol, tag):
    rb = np.load(os.path.join(G, f"reduced_basis_{tag}.npz"))
    out = so.greedy_reduced_basis(_dataset(name), tol)
    B, A = out[3], out[4]
    assert B.shape == rb["basis_reduced_ortho"].shape
    assert np.abs(B - rb["basis_reduced_ortho"]).max() < 1e-9            # unit vectors; deflation noise ~1e-11
    assert np.abs(A - rb["alpha_ortho"]).max() < 1e-9 * np.abs(rb["alpha_ortho"]).max()


def test_greedy_bookkeeping_outputs(ref):
    for tag, tol in (("g3", 1e-3), ("g6", 1e-6)):
        main_e, coef, pp, B, A, matP, nrm = so.greedy_reduced_basis(_dataset("BCC"), tol)
        assert (main_e == ref[f"{tag}_mainelem"]).all()
        np.testing.assert_allclose(coef, ref[f"{tag}_reducedcoef"], rtol=0, atol=1e-7 * np.abs(ref[f"{tag}_reducedcoef"]).max())
        np.testing.assert_allclose(matP, ref[f"{tag}_matP"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(nrm, ref[f"{tag}_norms"], rtol=1e-14)
        assert len(pp) == len(main_e)


def test_projection(ref):
    sd = _dataset("BCC")
    keys = list(sd)[:3]
    al = so.project_to_basis({k: sd[k] for k in keys}, ref["g6_basis"])
    np.testing.assert_allclose(np.stack([al[k] for k in keys]), ref["proj_alphas"], rtol=0, atol=1e-9 * np.abs(ref["proj_alphas"]).max())


def test_rbf_1d_and_reconstruction(ref):
    rb = np.load(os.path.join(G, "reduced_basis_BCC_tol_1e-6.npz"))
    x, a = rb["list_elements"], rb["alpha_ortho"].T
    wcp = so.tps_fit(x, a)
    al = so.tps_evaluate(x, wcp, ref["q1"])
    scale = np.abs(ref["s1_rbf_alphas"]).max()
    np.testing.assert_allclose(al, ref["s1_rbf_alphas"], rtol=0, atol=1e-10 * scale)
    S = so.schur_from_alphas(rb["basis_reduced_ortho"], al, 48)
    np.testing.assert_allclose(S, ref["s1_RBF"], rtol=0, atol=1e-10 * np.abs(ref["s1_RBF"]).max())
    g = so.tps_gradient(x, wcp, ref["q1"][:5])                           # (5, 1, 5)
    dS = np.stack([so.schur_from_alphas(rb["basis_reduced_ortho"], g[i], 48) for i in range(5)])
    np.testing.assert_allclose(dS, ref["s1_rbf_dS"], rtol=0, atol=1e-9 * np.abs(ref["s1_rbf_dS"]).max())


def test_linear_and_nearest(ref):
    rb = np.load(os.path.join(G, "reduced_basis_BCC_tol_1e-6.npz"))
    x, a = rb["list_elements"], rb["alpha_ortho"].T
    S = so.schur_from_alphas(rb["basis_reduced_ortho"], so.alphas_linear_1d(x, a, ref["q1"]), 48)
    np.testing.assert_allclose(S, ref["s1_linear"], rtol=0, atol=1e-12 * np.abs(S).max())
    S = so.schur_from_alphas(rb["basis_reduced_ortho"], so.alphas_nearest(x, a, ref["q1"]), 48)
    np.testing.assert_allclose(S, ref["s1_nearest_neighbor"], rtol=0, atol=1e-12 * np.abs(S).max())


def test_rbf_2d(ref):
    wcp = so.tps_fit(ref["x2"], ref["a2"])
    scale = np.abs(ref["r2_eval"]).max()
    np.testing.assert_allclose(so.tps_evaluate(ref["x2"], wcp, ref["q2"]), ref["r2_eval"], rtol=0, atol=1e-9 * scale)
    np.testing.assert_allclose(so.tps_gradient(ref["x2"], wcp, ref["q2"]), ref["r2_grad"], rtol=0,
                               atol=1e-9 * np.abs(ref["r2_grad"]).max())


def test_linear_nd(ref):
    np.testing.assert_allclose(so.alphas_linear_nd(ref["x2"], ref["a2"], ref["q2l"]), ref["l2_eval"], rtol=0,
                               atol=1e-12 * np.abs(ref["l2_eval"]).max())
