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
