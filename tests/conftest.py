import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_fixture(name):
    """A tests/golden/*.npz fixture as a dict (0-d arrays unwrapped)."""
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as data:
        return {k: (data[k].item() if data[k].ndim == 0 else data[k]) for k in data.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle import phylo
    phylo.load()
    return phylo
