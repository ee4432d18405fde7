"""CPU: host policy of the two-level preconditioner (aggregates, coarse inverse) -- torch on CPU tensors, no CUDA call."""
import numpy as np
import torch

from oracle import coarse_oracle as co
from pylatticedso_b200 import coarse
from pylatticedso_b200 import mesh as M


def test_box_aggregates_equal_the_oracle_rule():
    lat = M.synthetic_lattice("Octet", (5, 3, 4), [0.03])
    m = M.mesh_from_synthetic(lat, 1)
    for target in (1, 8, 27, 100):
        a_o, n_o = co.box_aggregates(m.xyz, target)
        a_d, n_d = coarse.box_aggregates(torch.from_numpy(m.x), torch.from_numpy(m.y), torch.from_numpy(m.z), target)
        assert n_d == n_o and (a_d.numpy() == a_o).all()
        assert np.bincount(a_o, minlength=n_o).min() > 0            # empty boxes are dropped


def test_default_aggregate_count():
    assert coarse.default_aggregates(100) == 8
    assert coarse.default_aggregates(265721) == 531               # Octet 40^3
    assert coarse.default_aggregates(4060301) == 1000             # Octet 100^3: capped


def test_invert_coarse_handles_dead_and_rank_deficient_blocks():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((12, 9))
    E = A @ A.T                                                     # rank 9 of 12: eigen-decomposition branch
    E[:, 3] = 0.0; E[3, :] = 0.0                                    # a fully constrained coarse DOF
    Einv = coarse.invert_coarse(torch.from_numpy(E)).numpy()
    assert np.abs(Einv - Einv.T).max() == 0.0
    assert np.abs(Einv[3]).max() == 0.0
    assert np.abs(E @ Einv @ E - E).max() < 1e-9 * np.abs(E).max()  # a generalised inverse on the range
    # comfortably positive definite: the Cholesky branch gives the plain inverse
    B = rng.standard_normal((8, 8))
    P = B @ B.T + 8 * np.eye(8)
    np.testing.assert_allclose(coarse.invert_coarse(torch.from_numpy(P)).numpy(), np.linalg.inv(P), rtol=0, atol=1e-12)


def test_stretch_dominated_rule():
    """The automatic choice of solve_FEM_B200(two_level="auto"): Octet (bulk valence 12) yes, BCC (8) no."""
    from pylatticedso_b200.fem import stretch_dominated
    assert stretch_dominated(M.mesh_from_synthetic(M.synthetic_lattice("Octet", (12, 12, 12), [0.03]), 1))
    assert stretch_dominated(M.mesh_from_synthetic(M.synthetic_lattice("Octet", (12, 12, 12), [0.03]), 3))   # subdivision does not matter
    assert not stretch_dominated(M.mesh_from_synthetic(M.synthetic_lattice("BCC", (12, 12, 12), [0.05]), 2))


def test_sharded_aggregates_agree_across_slabs():
    """What DistributedFEM.two_level does on the host: every rank cuts the boxes over the GLOBAL bounding box, so a node
    gets the same aggregate on the rank that owns it and on the ranks that see it as a ghost; compacting the global
    numbering gives the single-GPU (and oracle) aggregates."""
    from pylatticedso_b200 import distributed as D
    lat = M.synthetic_lattice("Octet", (6, 3, 3), [0.03])
    m = M.mesh_from_synthetic(lat, 1)
    lo, hi = m.xyz.min(0), m.xyz.max(0)
    nb, ext = coarse.box_grid(lo, hi, 12)
    t = torch.from_numpy
    g_idx = coarse.box_index(t(m.x), t(m.y), t(m.z), lo, ext, nb).numpy()
    assert g_idx.min() >= 0 and g_idx.max() < nb.prod()
    cen = coarse.box_centers(lo, ext, nb)
    assert cen.shape == (nb.prod(), 3) and np.all(cen >= lo) and np.all(cen <= hi)
    # every node lies in the box it is assigned to
    half = 0.5 * ext / nb
    assert np.all(np.abs(m.xyz - cen[g_idx]) <= half + 1e-12)
    world = 3
    owned_total = 0
    for rank in range(world):
        part = D.partition_slab(m, rank, world)
        lm = D.local_mesh(m, part)
        l_idx = coarse.box_index(t(lm.x), t(lm.y), t(lm.z), lo, ext, nb).numpy()      # local numbering [owned | ghosts]
        assert (l_idx == g_idx[part.local_nodes]).all()
        owned_total += part.n_owned
    assert owned_total == m.n_nodes                                                    # every node restricted exactly once
    a_o, n_o = co.box_aggregates(m.xyz, 12)
    uniq, inv = np.unique(g_idx, return_inverse=True)
    assert n_o == len(uniq) and (inv == a_o).all()
