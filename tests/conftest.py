import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


EMULATED = os.environ.get("MGB200_TEST_EMU") == "1"


def use_emulated_library():
    """MGB200_TEST_EMU=1 (set only by tests/test_emulated_library.py for its child pytest runs): point the ctypes
    binding at tests/host_emul/_build/libmgb200_emu.so -- the product's own csrc/ compiled with g++ against a
    CUDA-on-CPU emulation -- so that the `gpu`-marked parity tests can also run on a machine without a GPU.
    Test infrastructure only; the product never loads that library and the real `-m gpu` run never sets the variable."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "host_emul"))
    import build_emu
    import mgb200
    mgb200.capi.LIB_PATH = build_emu.build()
    mgb200.capi._lib = None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full-size BASELINE configs (GB of host memory, a minute or two); deselect with -m \"gpu and not slow\"")
    if EMULATED:
        use_emulated_library()


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
