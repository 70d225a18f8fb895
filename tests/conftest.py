import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "computational-fluid-dynamics_b200"))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def max_ulp(a, b):
    """Largest distance in units in the last place between two float64 arrays."""
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.int64).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.int64).astype(np.int64)
    a = np.where(a < 0, np.int64(-(2**63)) - a, a)
    b = np.where(b < 0, np.int64(-(2**63)) - b, b)
    return int(np.abs(a - b).max()) if a.size else 0


def rel_l2(a, b):
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(np.asarray(b).ravel())
    return d / n if n > 0 else d


@pytest.fixture(scope="session")
def orc():
    import orc as _orc
    _orc.orc_lib()
    return _orc


@pytest.fixture(scope="session")
def pm():
    import pm_ctypes
    pm_ctypes.lib()
    return pm_ctypes
