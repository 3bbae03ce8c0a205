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


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle import prism_oracle
    prism_oracle.lib()          # builds liboracle_prism.so if missing
    return prism_oracle


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a machine that has a GPU (fails loudly otherwise)."""
    from fatiando_b200 import _lib
    _lib.require_device()
    return _lib
