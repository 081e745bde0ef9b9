import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure; never used by the product)."""
    import oracle
    return oracle.get()


@pytest.fixture(scope="session")
def mgb():
    import mgb200
    return mgb200


def rand_vec(level, dtype, seed, scale=1.0):
    n = (1 << level) - 1
    rng = np.random.default_rng(seed)
    return (scale * rng.uniform(-1.0, 1.0, n * n)).astype(dtype)


def assert_bitwise(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, f"{what}: shape/dtype {a.shape}/{a.dtype} vs {b.shape}/{b.dtype}"
    if not np.array_equal(a, b):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        k = int(np.argmax(d))
        denom = max(float(np.abs(b).max()), 1e-300)
        raise AssertionError(f"{what}: not bitwise equal; max|diff|={d.max():.3e} (rel {d.max() / denom:.3e}) "
                             f"at {k}: {a.flat[k]!r} vs {b.flat[k]!r}; mismatches={int((a != b).sum())}/{a.size}")
