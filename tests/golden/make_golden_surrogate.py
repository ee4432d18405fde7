"""Freeze outputs of the reference's OWN reduced-basis / surrogate code.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_surrogate.py

Writes
* tests/golden/reduced_basis_*.npz  -- copies of the reference's stored reduced bases (reference-held vectors),
* tests/golden/surrogate_ref.npz    -- outputs of reduce_basis_greedy (greedy_algorithm.py:35-155),
  ThinPlateSplineRBF (utils_rbf.py), LatticeSim.get_schur_complement_from_reduced_basis_batch (lattice_sim.py:921-978),
  _compute_schur_gradients_RBF (:1056-1082), evaluate_alphas_linear_surrogate (:755-807) and the nearest-neighbour
  look-up (:939-942), all called unmodified on the reference's stored data.
"""
import importlib.util
import os
import shutil
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from pylatticedso_b200 import refshim
    ls = refshim.import_reference()
    greedy = _load(os.path.join(REF, "src/pyLatticeSim/greedy_algorithm.py"), "ref_greedy")
    rbf = _load(os.path.join(REF, "src/pyLatticeSim/utils_rbf.py"), "ref_rbf")
    rb_dir = os.path.join(REF, "data/outputs/schur_complement/reduced_basis")
    out = {}
    for name in ("BCC_tol_1e-3", "BCC_tol_1e-6", "Hybrid1_tol_1e-6", "Hybrid4_tol_1e-6"):
        shutil.copyfile(os.path.join(rb_dir, f"reduced_basis_{name}.npz"), os.path.join(HERE, f"reduced_basis_{name}.npz"))

    # --- greedy: every return value of the reference function on its own BCC dataset
    d = np.load(os.path.join(REF, "data/outputs/schur_complement/Schur_complement_BCC.npz"))
    sd = {tuple(r): S for r, S in zip(d["radius_values"], d["schur_matrices"])}
    for tag, tol in (("g3", 1e-3), ("g6", 1e-6)):
        main_e, coef, _pp, B, A, matP, nrm = greedy.reduce_basis_greedy(sd, tol, verbose=0)
        out.update({f"{tag}_mainelem": main_e, f"{tag}_reducedcoef": coef, f"{tag}_basis": B, f"{tag}_alpha": A,
                    f"{tag}_matP": matP, f"{tag}_norms": nrm})
    # projection (C-order ravel) of three snapshots on the tol-1e-6 basis
    proj = greedy.project_to_reduced_basis({k: sd[k] for k in list(sd)[:3]}, out["g6_basis"])
    out["proj_alphas"] = np.stack([proj[k] for k in list(sd)[:3]])

    # --- 1-D surrogate (BCC, tol 1e-6): RBF / linear / nearest through the reference's LatticeSim methods
    rb = np.load(os.path.join(rb_dir, "reduced_basis_BCC_tol_1e-6.npz"))
    rng = np.random.default_rng(7)
    q1 = np.concatenate([rng.uniform(0.008, 0.105, 13), [0.01, 0.05, 0.1]])[:, None]      # incl. centres (r = 0) and outside
    for kind in ("RBF", "linear", "nearest_neighbor"):
        me = types.SimpleNamespace(type_schur_complement_computation=kind, reduce_basis_dict=rb,
                                   alpha_coefficients_greedy=rb["alpha_ortho"].T, radial_basis_function=None,
                                   shape_schur_complement=None, _verbose=0)
        me._define_radial_basis_functions = types.MethodType(ls.LatticeSim._define_radial_basis_functions, me)
        f = ls.LatticeSim.evaluate_alphas_linear_surrogate
        me.evaluate_alphas_linear_surrogate = types.MethodType(getattr(f, "__wrapped__", f), me)
        if kind == "nearest_neighbor":
            from sklearn.neighbors import NearestNeighbors
            me.neigh_function = NearestNeighbors(n_neighbors=1, algorithm="auto").fit(rb["list_elements"])
        S = ls.LatticeSim.get_schur_complement_from_reduced_basis_batch(me, [list(x) for x in q1])
        out[f"s1_{kind}"] = S
        if kind == "RBF":
            out["s1_rbf_W"] = me.radial_basis_function.W
            out["s1_rbf_CP"] = me.radial_basis_function.CP
            out["s1_rbf_alphas"] = me.radial_basis_function.evaluate(q1)
            out["s1_rbf_dS"] = np.stack([np.stack(ls.LatticeSim._compute_schur_gradients_RBF(me, list(x))) for x in q1[:5]])
    out["q1"] = q1

    # --- 2-D RBF (BCC + Hybrid4, 100 centres, 38 coefficients): the interpolant only (the basis is 1.6 MB)
    rb2 = np.load(os.path.join(rb_dir, "reduced_basis_BCC_Hybrid4_tol_1e-6.npz"))
    x2, a2 = rb2["list_elements"], rb2["alpha_ortho"].T
    q2 = np.vstack([rng.uniform(x2.min(0), x2.max(0), (20, 2)), x2[[0, 17, 99]]])
    r2 = rbf.ThinPlateSplineRBF(x2, a2)
    out.update({"x2": x2, "a2": a2, "q2": q2, "r2_eval": r2.evaluate(q2), "r2_grad": r2.gradient(q2)})
    # --- 2-D "linear" surrogate (scipy Delaunay interpolation inside the hull, nearest neighbour outside,
    #     lattice_sim.py:794-807) through the reference's own method, incl. queries outside the convex hull
    q2l = np.vstack([q2, x2.min(0) - 0.01, x2.max(0) + 0.02, [x2[:, 0].mean(), x2[:, 1].max() + 0.05]])
    me = types.SimpleNamespace(reduce_basis_dict={"list_elements": x2}, alpha_coefficients_greedy=a2, _verbose=0)
    f = ls.LatticeSim.evaluate_alphas_linear_surrogate
    f = getattr(f, "__wrapped__", f)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        out["l2_eval"] = np.stack([np.asarray(f(me, list(q))) for q in q2l])
    out["q2l"] = q2l
    np.savez_compressed(os.path.join(HERE, "surrogate_ref.npz"), **out)
    print({k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
