"""CPU check of the shared-memory TILE kernels (csrc/tile_core.h) by exact host emulation.

tests/host_emul/tile_emul.cpp compiles the very phase functions the CUDA kernel `k_tile` runs
(plain C++, no CUDA) and executes them CTA by CTA, phase by phase, thread by thread.  The
results must equal the oracle's composition of the unfused operators bit for bit -- the same
contract the GPU parity tests enforce for the other kernels.  (The tile kernels were written in
round 1 after the GPU budget was spent; this is how they were verified.  They are selected at
run time only with MGB200_TILE=1 until the GPU parity tests have covered them.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import assert_bitwise, rand_vec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "multigrid_nikhil_c-_b200", "csrc")
SWEEPS, PRE, POST = 0, 1, 2


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libtile_emul.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I" + CSRC,
                    os.path.join(ROOT, "tests", "host_emul", "tile_emul.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def pad(vec, level, dtype):
    """interior n*n vector -> padded node grid (N+1 rows, pitch = round_up(N+1, 32)), zero ring"""
    N = 1 << level
    n = N - 1
    pitch = (N + 1 + 31) // 32 * 32
    a = np.zeros((N + 1, pitch), dtype=dtype)
    a[1:N, 1:N] = np.asarray(vec, dtype=dtype).reshape(n, n)
    return a


def unpad(a, level):
    N = 1 << level
    return np.ascontiguousarray(a[1:N, 1:N]).reshape(-1)


def run(emul, mode, ns, rbgs, tile, level, x, b, e=None, dtype=np.float64, nthr=256, rows=None, omega=2.0 / 3.0):
    N = 1 << level
    U, F = pad(x, level, dtype), pad(b, level, dtype)
    out = np.full_like(U, np.nan)
    pitch = U.shape[1]
    ya, yb = rows if rows else (1, N)
    Nc = N // 2
    pitch_c = (Nc + 1 + 31) // 32 * 32
    fc = np.zeros((Nc + 1, pitch_c), dtype=dtype)
    uc = np.full((Nc + 1, pitch_c), 7.0, dtype=dtype)
    ec = pad(e, level - 1, dtype) if e is not None else np.zeros((Nc + 1, pitch_c), dtype=dtype)
    c0, c1 = oracle.get().jacobi_constants(omega, dtype)
    fn = emul.tile_emul_f64 if np.dtype(dtype) == np.float64 else emul.tile_emul_f32
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = fn(mode, ns, int(rbgs), tile, nthr, N, ctypes.c_longlong(pitch), ya, yb, 0, N + 1, P(U), P(out), P(F),
            ctypes.c_double(c0), ctypes.c_double(c1), ctypes.c_double(0.25), P(fc), P(uc), P(ec),
            ctypes.c_longlong(pitch_c), 0, Nc + 1)
    assert rc == 0
    return out, fc, uc


def smooth(o, x, b, ns, rbgs):
    return o.rbgs(x, b, ns // 2) if rbgs else o.jacobirelaxation(x, b, ns)


CASES = [(np.float64, 0), (np.float64, 1), (np.float32, 0), (np.float32, 1)]


@pytest.mark.parametrize("dtype,tile", CASES)
@pytest.mark.parametrize("level", [2, 3, 5, 7])
def test_tile_sweeps(emul, orc, level, dtype, tile):
    x, b = rand_vec(level, dtype, 61), rand_vec(level, dtype, 62, 1e-3)
    for rbgs, nss in ((False, (1, 2, 3)), (True, (2, 4))):
        for ns in nss:
            out, _, _ = run(emul, SWEEPS, ns, rbgs, tile, level, x, b, dtype=dtype)
            N = 1 << level
            assert not out[1:N, 0].any() and not np.isnan(out[1:N, :N]).any()      # ring column stays zero
            assert_bitwise(unpad(out, level), smooth(orc, x, b, ns, rbgs), f"tile sweeps ns={ns} rbgs={rbgs} L{level}")


@pytest.mark.parametrize("dtype,tile", CASES)
@pytest.mark.parametrize("level", [2, 3, 5, 7])
def test_tile_pre(emul, orc, level, dtype, tile):
    """PRE: NS sweeps, residual, full weighting (restriction2d P:531-546), zero coarse guess (P:613)."""
    x, b = rand_vec(level, dtype, 63), rand_vec(level, dtype, 64, 1e-3)
    for rbgs, nss in ((False, (1, 2)), (True, (2, 4))):
        for ns in nss:
            out, fc, uc = run(emul, PRE, ns, rbgs, tile, level, x, b, dtype=dtype)
            u = smooth(orc, x, b, ns, rbgs)
            assert_bitwise(unpad(out, level), u, f"tile pre u ns={ns} rbgs={rbgs} L{level}")
            assert_bitwise(unpad(fc, level - 1), orc.restriction2d(orc.residual(u, b)), f"tile pre f_c ns={ns} rbgs={rbgs}")
            Nc = 1 << (level - 1)
            assert not uc[1:Nc, 1:Nc].any()                       # zero guess written on the coarse interior
            assert not fc[0].any() and not fc[:, 0].any() and not fc[Nc].any() and not fc[:, Nc:].any()


@pytest.mark.parametrize("dtype,tile", CASES)
@pytest.mark.parametrize("level", [2, 3, 5, 7])
def test_tile_post(emul, orc, level, dtype, tile):
    """POST: bilinear prolongation + correction (P:337-425, P:620-624), then NS sweeps."""
    x, b, e = rand_vec(level, dtype, 65), rand_vec(level, dtype, 66, 1e-3), rand_vec(level - 1, dtype, 67)
    for rbgs, nss in ((False, (1, 2)), (True, (2, 4))):
        for ns in nss:
            out, _, _ = run(emul, POST, ns, rbgs, tile, level, x, b, e=e, dtype=dtype)
            want = smooth(orc, orc.prolong_correct(e, x), b, ns, rbgs)
            assert_bitwise(unpad(out, level), want, f"tile post ns={ns} rbgs={rbgs} L{level}")


def test_tile_thread_count_and_row_ranges_do_not_matter(emul, orc):
    """Any block size gives the same bits; a sub-range of rows (a slab) writes exactly those rows."""
    level, dtype = 6, np.float64
    x, b = rand_vec(level, dtype, 68), rand_vec(level, dtype, 69, 1e-3)
    want = orc.jacobirelaxation(x, b, 2)
    n = (1 << level) - 1
    for nthr in (32, 96, 256):
        out, _, _ = run(emul, SWEEPS, 2, False, 0, level, x, b, nthr=nthr)
        assert_bitwise(unpad(out, level), want, f"nthr={nthr}")
    out, _, _ = run(emul, SWEEPS, 2, False, 1, level, x, b, rows=(17, 40))
    got = out[1:n + 1, 1:n + 1]
    assert np.array_equal(got[16:39], want.reshape(n, n)[16:39])
    assert np.isnan(got[:16]).all() and np.isnan(got[39:]).all()
