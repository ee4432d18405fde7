"""CPU: host-side flattening and the vectorised generator against the reference numbering."""
import numpy as np
import pytest

from conftest import load_golden
from pylatticedso_b200 import mesh as M
from pylatticedso_b200 import refshim


@pytest.mark.parametrize("name", ["bcc_322", "octet_322", "octet_223_graded", "bcc_345"])
def test_synthetic_lattice_reproduces_reference_numbering(name):
    G = load_golden(f"numbering_{name}.npz")
    grad = None
    if float(G["grad_param_z"]) != 0.0:
        grad = ("linear", [False, False, True], [0.0, 0.0, float(G["grad_param_z"])])
    lat = M.synthetic_lattice(str(G["geom"]), tuple(G["n_cells"]), list(G["radii"]), grad_radius=grad)
    assert np.array_equal(lat.pxyz, G["pxyz"])          # node.index order + coordinates, bit exact
    assert np.array_equal(lat.b_p1, G["b_p1"]) and np.array_equal(lat.b_p2, G["b_p2"])
    assert np.array_equal(lat.b_rad, G["b_rad"])        # shared struts keep the first cell's radius
    assert np.array_equal(lat.b_cell, G["b_cell"])
    assert np.array_equal(lat.cell_radii, G["cell_radii"])


def test_node_and_strut_counts_match_closed_forms():
    for n in (3, 5):
        b = M.synthetic_lattice("BCC", (n, n, n), [0.05])
        assert b.pxyz.shape[0] == (n + 1) ** 3 + n ** 3 and b.b_p1.shape[0] == 8 * n ** 3
        o = M.synthetic_lattice("Octet", (n, n, n), [0.03])
        assert o.pxyz.shape[0] == (n + 1) ** 3 + 3 * n * n * (n + 1)
        assert o.b_p1.shape[0] == 12 * n * n * (n + 1) + 12 * n ** 3


def test_subdivision_layout():
    lat = M.synthetic_lattice("BCC", (2, 1, 1), [0.05])
    m = M.mesh_from_synthetic(lat, 3)
    nb = lat.b_p1.shape[0]
    assert m.n_elems == 3 * nb and m.n_nodes == lat.pxyz.shape[0] + 2 * nb
    # element chain of beam 0: p1 -> i0 -> i1 -> p2, interior nodes appended beam-major
    npnt = lat.pxyz.shape[0]
    assert list(m.en0[:3]) == [lat.b_p1[0], npnt, npnt + 1] and list(m.en1[:3]) == [npnt, npnt + 1, lat.b_p2[0]]
    a, c = lat.pxyz[lat.b_p1[0]], lat.pxyz[lat.b_p2[0]]
    assert np.allclose(m.xyz[npnt], a + (c - a) / 3)
    assert M.gmsh_segments(np.array([0.8660254]), 0.05)[0] == 18   # BCC half diagonal, h = 0.05
    assert M.gmsh_segments(np.array([0.01]), 0.05)[0] == 1


def test_compression_bc_counts():
    lat = M.synthetic_lattice("BCC", (5, 5, 5), [0.05])
    m = M.mesh_from_synthetic(lat, 1)
    fixed, g, f = M.compression_bc(m)
    assert int(fixed.sum()) == 252                       # SURVEY.md section 8d, C1 [probed on the reference]
    assert m.n_nodes == 341 and m.n_elems == 1000 and m.n_dof == 2046
    assert (g[fixed == 1] != 0).sum() == 36 and not f.any()


@pytest.mark.skipif(not refshim.have_reference(), reason="needs a pyLatticeDSO checkout")
def test_flatten_object_graph_matches_generator_and_oracle():
    import contextlib
    import io
    from oracle import lattice_oracle as orc
    ls = refshim.import_reference()
    cfg = {"geometry": {"cell_size": {"x": 1, "y": 1, "z": 1}, "number_of_cells": {"x": 2, "y": 3, "z": 2},
                        "radii": [0.04], "geom_types": ["Octet"]},
           "simulation_parameters": {"enable": False, "material": "VeroClear", "periodicity": False},
           "boundary_conditions": {"Displacement": {"Fixed": {"Surface": ["Zmin"], "DOF": ["X", "Y", "Z", "RX", "RY", "RZ"],
                                                              "Value": [0, 0, 0, 0, 0, 0]},
                                                    "Load": {"Surface": ["Zmax"], "DOF": ["Z"], "Value": [-0.01]}}}}
    refshim.set_inline_presets({"t": cfg})
    with contextlib.redirect_stdout(io.StringIO()):
        lat = ls.LatticeSim("t")
    syn = M.synthetic_lattice("Octet", (2, 3, 2), [0.04])
    for m_ in (1, 2, "gmsh"):
        a = M.flatten_lattice(lat, None, m_)
        b = M.mesh_from_synthetic(syn, m_)
        o = orc.flatten_lattice(lat, None, m_)
        for k in ("x", "y", "z", "en0", "en1", "rad"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (m_, k)
        assert np.array_equal(a.xyz, o["xyz"]) and np.array_equal(a.en0, o["en"][:, 0]) and np.array_equal(a.rad, o["rad"])
    a = M.flatten_lattice(lat, None, 1)
    fixed, g, f = M.bc_arrays_from_lattice(lat, a)
    fs, gs, fs_ = M.compression_bc(M.mesh_from_synthetic(syn, 1))
    assert np.array_equal(fixed, fs) and np.array_equal(g, gs) and np.array_equal(f, fs_)


@pytest.mark.parametrize("geom", ["BCC", "Hybrid1", "Hybrid4"])
def test_strut_chains_of_the_golden_cells(geom):
    """Host side of the Schur strut pre-pass: every element lies in exactly one chain, chains run joint to joint,
    boundary nodes keep their index in the reduced numbering, and the chain set equals the oracle's."""
    from conftest import load_golden, mesh_from_npz
    from oracle import lattice_oracle as orc
    from pylatticedso_b200.schur import local_cell_mesh, strut_chains
    G = load_golden(f"schur_{geom}.npz")
    m = mesh_from_npz(G, "c0_")
    bnd = G["c0_bnd"].reshape(-1, 6)[:, 0] // 6
    perm, xyz, l0, l1 = local_cell_mesh(m, bnd)
    ch = strut_chains(xyz, l0, l1, len(bnd))
    assert ch is not None and ch["ptr"][0] == 0 and ch["ptr"][-1] == len(l0)
    assert sorted(ch["elem"].tolist()) == list(range(len(l0)))
    assert ch["a"].max() < ch["n_joints"] and ch["b"].max() < ch["n_joints"] and ch["n_joints"] >= len(bnd)
    oc = orc.find_chains(xyz.shape[0], np.stack([l0, l1], 1), np.arange(len(bnd)), xyz)
    assert len(oc) == len(ch["a"])
    assert sorted(tuple(sorted(e for e, _ in c[2])) for c in oc) == sorted(
        tuple(sorted(ch["elem"][ch["ptr"][k]:ch["ptr"][k + 1]].tolist())) for k in range(len(ch["a"])))
    # walking order: consecutive elements share a node, flips orient them from a to b
    for k in range(len(ch["a"])):
        prev = None
        for q in range(ch["ptr"][k], ch["ptr"][k + 1]):
            e, fl = ch["elem"][q], ch["flip"][q]
            a, b = (l1[e], l0[e]) if fl else (l0[e], l1[e])
            assert prev is None or prev == a
            prev = b


def test_strut_chains_none_for_single_element_struts():
    from pylatticedso_b200 import mesh as M
    from pylatticedso_b200.schur import bcc_cell_order_nodes, local_cell_mesh, strut_chains
    lat = M.synthetic_lattice("BCC", (1, 1, 1), [0.05])
    mesh = M.mesh_from_synthetic(lat, 1)
    bnd = bcc_cell_order_nodes(lat.pxyz, (0, 1, 0, 1, 0, 1))
    perm, xyz, l0, l1 = local_cell_mesh(mesh, bnd)
    assert strut_chains(xyz, l0, l1, len(bnd)) is None



@pytest.mark.parametrize("geom,n,radii,kw", [
    ("BCC", (3, 2, 2), [0.05], {}),
    ("Octet", (3, 4, 2), [0.03], {}),
    (["BCC", "Octet"], (2, 3, 3), [0.05, 0.02], {}),
    ("Octet", (6, 3, 2), [0.03], dict(i_range=(2, 5))),
    ("BCC", (5, 3, 4), [0.04], dict(grad_radius=("linear", [1, 0, 1], [0.1, 0, 0.2]))),
    ("BCC", (4, 2, 3), [0.04], dict(cell_size=(0.1, 0.3, 0.7))),
    ("Octet", (3, 3, 3), [0.04], dict(cell_radii=np.random.default_rng(0).uniform(0.01, 0.05, (27, 1)))),
])
def test_sort_free_grid_generator_is_bit_identical_to_the_generic_numbering(geom, n, radii, kw):
    """mesh._grid_lattice (dense scatters + prefix sums on the half-cell grid) against the sort/unique restatement of
    the reference's numbering rules, which the numbering_*.npz fixtures pin to the reference objects."""
    from pylatticedso_b200 import mesh as M
    a = M.synthetic_lattice(geom, n, radii, **kw)
    b = M.synthetic_lattice(geom, n, radii, _force_generic=True, **kw)
    for f in ("pxyz", "b_p1", "b_p2", "b_rad", "b_cell", "b_type", "cell_radii"):
        x, y = getattr(a, f), getattr(b, f)
        assert x.shape == y.shape and np.array_equal(x, y), f
    ma, mb = M.mesh_from_synthetic(a, 1), M.mesh_from_synthetic(b, 2)
    assert ma.n_elems == a.b_p1.shape[0] and mb.n_elems == 2 * ma.n_elems
    np.testing.assert_array_equal(ma.en0, a.b_p1)


def test_grid_generator_on_torch_tensors_is_bit_identical():
    """mesh._grid_lattice_torch (the device form of the half-cell-grid numbering; scatter_reduce(amin) = "first creation
    wins") against the numpy path, on CPU tensors: full lattices, a slab range, two geometries, graded and per-cell radii."""
    import numpy as np
    from pylatticedso_b200 import mesh as M
    cases = [("BCC", (5, 4, 3), [0.05], {}), ("Octet", (6, 5, 4), [0.03], {}), ("Octet", (9, 4, 4), [0.03], {"i_range": (3, 6)}),
             (["BCC", "Octet"], (4, 3, 3), [0.05, 0.02], {}),
             ("Octet", (5, 5, 5), [0.03], {"grad_radius": ("linear", [False, False, True], [0, 0, 0.1])}),
             ("BCC", (4, 4, 4), [1.0], {"cell_radii": np.random.default_rng(0).random((64, 1))})]
    for geom, n, r, kw in cases:
        a = M.synthetic_lattice(geom, n, r, **kw)
        b = M.synthetic_lattice(geom, n, r, device="cpu", **kw)
        for k in ("pxyz", "b_p1", "b_p2", "b_rad", "b_cell", "b_type"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (geom, n, k)
            assert getattr(a, k).dtype == getattr(b, k).dtype
