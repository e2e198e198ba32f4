import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _build_if_missing():
    """The built libraries are not in git (they travel to the GPU box with the
    snapshot): in a fresh checkout build the CUDA library and the oracle once, the
    way __graft_entry__.build() does, so that the suite can run at all."""
    import shutil
    import subprocess
    product = os.path.join(ROOT, "libsbn_b200", "lib", "libsbn_b200.so")
    checker = os.path.join(ROOT, "oracle", "_build", "libphylo_oracle.so")
    if not os.path.exists(product) and shutil.which("nvcc"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "libsbn_b200", "csrc"), "-j8"], check=True,
                       stdout=subprocess.DEVNULL)
    if not os.path.exists(checker) and shutil.which("g++"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True,
                       stdout=subprocess.DEVNULL)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    _build_if_missing()


def load_fixture(name):
    """A tests/golden/*.npz fixture as a dict (0-d arrays unwrapped)."""
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as data:
        return {k: (data[k].item() if data[k].ndim == 0 else data[k]) for k in data.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle import phylo
    phylo.load()
    return phylo
