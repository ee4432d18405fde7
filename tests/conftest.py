import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
E_MOD, NU = 1013.0, 0.3  # VeroClear


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def built_lib():
    from pylatticedso_b200 import lib
    lib.build()
    return lib.load()


@pytest.fixture(scope="session")
def ctx(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pylatticedso_b200 import lib
    c = lib.Context()
    yield c
    c.close()


def mesh_from_npz(G, prefix=""):
    from pylatticedso_b200.mesh import BeamMesh
    return BeamMesh(x=G[prefix + "x"], y=G[prefix + "y"], z=G[prefix + "z"], en0=G[prefix + "en0"],
                    en1=G[prefix + "en1"], rad=G[prefix + "rad"], beam_of_elem=G[prefix + "beam_of_elem"],
                    chain=G[prefix + "chain"], n_points=int(G[prefix + "n_points"]),
                    point_index=G[prefix + "point_index"], cell_of_elem=G[prefix + "cell_of_elem"])
