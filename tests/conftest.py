import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def relerr_mat(x, ref):
    """SURVEY §8c parity metric for matrices: max|x - ref| / max|ref|."""
    x = np.asarray(x, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    den = float(np.max(np.abs(ref)))
    return float(np.max(np.abs(x - ref))) / (den if den > 0 else 1.0)


def relerr_vec(x, ref):
    """element-wise relative error for positive vectors."""
    x = np.asarray(x, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    return float(np.max(np.abs(x - ref) / np.abs(ref)))


TOL = 1e-10   # BASELINE.json north_star: 1e-10 relative in FP64


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "fable_golden.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def ctx():
    import fable_b200
    c = fable_b200.Context(0, 12345)
    yield c
    c.close()
