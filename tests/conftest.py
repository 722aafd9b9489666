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


@pytest.fixture(scope="session")
def emor():
    z = np.load(os.path.join(GOLDEN, "invemor_f32.npz"))
    return z["B"], z["g0"], z["hinv"]


@pytest.fixture(scope="session")
def golden_small():
    return dict(np.load(os.path.join(GOLDEN, "oracle_small.npz")))


@pytest.fixture(scope="session")
def kat_lin2():
    return dict(np.load(os.path.join(GOLDEN, "kat_lin2.npz")))


@pytest.fixture(scope="session")
def shdr_gpu(emor):
    """The product package with the EMoR table installed; fails (not skips) without the .so."""
    import shdr
    shdr.require_gpu()
    shdr.set_emor_table(emor[1], emor[2])
    return shdr
