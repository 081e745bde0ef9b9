"""CPU check of the CLUSTER coarse tail (csrc/ctail_core.h) by exact host emulation.

tests/host_emul/ctail_emul.cpp executes the op list of `k_ctail` -- the same host-device functions, one op per
cluster barrier -- for every CTA of the cluster and every thread, with "distributed shared memory" being the
other CTAs' arrays.  One launch must reproduce a whole V- or W-cycle of the oracle from level <= 8 down to the
coarsest level and back, bit for bit, for any cluster size.  (Written in round 1 after the GPU budget was
spent; opt-in at run time with MGB200_CTAIL=1 until a GPU has run it.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import assert_bitwise, rand_vec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "multigrid_nikhil_c-_b200", "csrc")


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ctail") / "libctail_emul.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I" + CSRC,
                    os.path.join(ROOT, "tests", "host_emul", "ctail_emul.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    lib.ctail_smem_bytes_f64.restype = ctypes.c_longlong
    lib.ctail_smem_bytes_f32.restype = ctypes.c_longlong
    return lib


def run(emul, level, x, b, p: oracle.Params, nctas, dtype=np.float64, nthr=64):
    N = 1 << level
    n = N - 1
    pitch = (N + 1 + 31) // 32 * 32
    U = np.zeros((N + 1, pitch), dtype=dtype)
    F = np.zeros((N + 1, pitch), dtype=dtype)
    U[1:N, 1:N] = x.reshape(n, n)
    F[1:N, 1:N] = b.reshape(n, n)
    c0, c1 = oracle.get().jacobi_constants(p.omega, dtype)
    fn = emul.ctail_emul_f64 if np.dtype(dtype) == np.float64 else emul.ctail_emul_f32
    nops = ctypes.c_int()
    rc = fn(level, p.coarsest_level, p.nu1, p.nu2, p.gamma, p.smoother, nctas, nthr, ctypes.c_longlong(pitch),
            U.ctypes.data_as(ctypes.c_void_p), F.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(c0), ctypes.c_double(c1),
            ctypes.c_double(p.restrict_weight), ctypes.byref(nops))
    assert rc == 0
    assert not U[0].any() and not U[N].any() and not U[:, 0].any() and not U[:, N:].any()     # ring untouched
    return np.ascontiguousarray(U[1:N, 1:N]).reshape(-1), nops.value


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nctas", [1, 2, 4, 16])
@pytest.mark.parametrize("level,coarsest", [(1, 1), (3, 1), (5, 2), (6, 1), (7, 1)])
def test_cluster_tail_vcycle_bitwise(emul, orc, level, coarsest, nctas, dtype):
    x, b = rand_vec(level, dtype, 91), rand_vec(level, dtype, 92, 1e-3)
    for smoother, nu1, nu2, gamma in ((0, 2, 2, 1), (0, 1, 3, 2), (1, 2, 2, 1), (1, 1, 1, 2)):
        p = oracle.Params(coarsest_level=coarsest, nu1=nu1, nu2=nu2, gamma=gamma, smoother=smoother)
        got, nops = run(emul, level, x, b, p, nctas, dtype)
        assert_bitwise(got, orc.vcyclemultigrid(x, b, p), f"ctail L{level} C={nctas} s={smoother} nu=({nu1},{nu2}) g={gamma}")


def test_cluster_tail_level8_16_ctas(emul, orc):
    """The target configuration: levels <= 8 (257^2) on a 16-CTA cluster; shared memory must fit 227 KB per CTA."""
    level = 8
    assert emul.ctail_smem_bytes_f64(8, 16) <= 227 * 1024
    assert emul.ctail_smem_bytes_f64(7, 4) <= 227 * 1024
    assert emul.ctail_smem_bytes_f32(8, 8) <= 227 * 1024
    x, b = rand_vec(level, np.float64, 93), rand_vec(level, np.float64, 94, 1e-3)
    for p in (oracle.Params(), oracle.Params(gamma=2), oracle.Params(smoother=1)):
        got, nops = run(emul, level, x, b, p, 16)
        assert_bitwise(got, orc.vcyclemultigrid(x, b, p), f"ctail L8 C=16 {p}")
    # a V(2,2) over 8 levels is 7 ops per level visit (+ load/store)
    _, nops = run(emul, level, x, b, oracle.Params(), 16)
    assert nops == 2 + 7 * 7 + 4


def test_cluster_tail_thread_count_does_not_matter(emul, orc):
    level = 6
    x, b = rand_vec(level, np.float64, 95), rand_vec(level, np.float64, 96, 1e-3)
    want = orc.vcyclemultigrid(x, b, oracle.Params())
    for nthr in (1, 32, 512):
        got, _ = run(emul, level, x, b, oracle.Params(), 8, nthr=nthr)
        assert_bitwise(got, want, f"nthr={nthr}")
