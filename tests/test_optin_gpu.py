"""GPU parity tests of the OPT-IN code paths (written in round 1 after the GPU budget was spent, CPU-verified
only): shared-memory tile kernels (MGB200_TILE=1) and, under torchrun via tests/mgpu_worker.py, the
communication-avoiding slab schedule (MGB200_COMM_AVOID=1) and distributed graph capture (MGB200_GRAPH_DIST=1).

They are skipped unless MGB200_TEST_OPTIN=1, so that an unverified path can never turn the default GPU suite red:
    MGB200_TEST_OPTIN=1 python -m pytest tests/test_optin_gpu.py -m gpu -x -q
The knobs are read from the environment when a context is created, so they are toggled in-process."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from conftest import assert_bitwise, rand_vec

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MGB200_TEST_OPTIN") != "1", reason="opt-in paths: set MGB200_TEST_OPTIN=1")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def knob():
    saved = {}

    def set_knob(name, value):
        saved.setdefault(name, os.environ.get(name))
        os.environ[name] = value
    yield set_knob
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("smoother,nu1,nu2,gamma", [("jacobi", 2, 2, 1), ("jacobi", 1, 1, 1), ("jacobi", 2, 1, 2),
                                                    ("rbgs", 2, 2, 1), ("rbgs", 1, 1, 2), ("jacobi", 4, 3, 1)])
@pytest.mark.parametrize("level", [3, 5, 7, 8, 10])
def test_tile_kernels_cycles_bitwise(mgb, orc, knob, level, dtype, smoother, nu1, nu2, gamma):
    knob("MGB200_TILE", "1")
    x, b = rand_vec(level, dtype, 81), rand_vec(level, dtype, 82, 1e-3)
    p = oracle.Params(nu1=nu1, nu2=nu2, gamma=gamma, smoother=1 if smoother == "rbgs" else 0, nthreads=4)
    want = [x]
    for _ in range(3):
        want.append(orc.vcyclemultigrid(want[-1], b, p))
    for graph, tail in ((False, False), (True, True)):
        with mgb.Multigrid(level, dtype=dtype, smoother=smoother, graph=graph, fused=True, coarse_tail=tail) as mg:
            mg.set_u(level, x)
            mg.set_rhs(level, b)
            for k in range(3):
                mg.cycle(level, nu1, nu2, gamma)
                assert_bitwise(mg.get_u(level), want[k + 1], f"tile cycle {k + 1} graph={graph} tail={tail}")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("smoother,nu1,nu2,gamma", [("jacobi", 2, 2, 1), ("jacobi", 1, 2, 2), ("rbgs", 2, 2, 1), ("rbgs", 1, 1, 2),
                                                    ("jacobi", 3, 2, 1)])
@pytest.mark.parametrize("level", [4, 7, 8, 10])
def test_zero_guess_chain_bitwise(mgb, orc, knob, level, dtype, smoother, nu1, nu2, gamma):
    """MGB200_ZERO_GUESS=1: PRE skips the zero coarse guess store, the next level's PRE / the tail do not read u."""
    knob("MGB200_ZERO_GUESS", "1")
    x, b = rand_vec(level, dtype, 83), rand_vec(level, dtype, 84, 1e-3)
    p = oracle.Params(nu1=nu1, nu2=nu2, gamma=gamma, smoother=1 if smoother == "rbgs" else 0, nthreads=4)
    want = [x]
    for _ in range(3):
        want.append(orc.vcyclemultigrid(want[-1], b, p))
    for graph, tail in ((False, False), (False, True), (True, True)):
        with mgb.Multigrid(level, dtype=dtype, smoother=smoother, graph=graph, fused=True, coarse_tail=tail) as mg:
            mg.set_u(level, x)
            mg.set_rhs(level, b)
            for k in range(3):
                mg.cycle(level, nu1, nu2, gamma)
                assert_bitwise(mg.get_u(level), want[k + 1], f"zero-guess cycle {k + 1} graph={graph} tail={tail}")
            # every API that reads a coarse iterate must see real zeros
            if level > 2:
                mg.residual(level)
                mg.restrict(level)
                assert not mg.get_u(level - 1).any()
            pv = oracle.Params(nu1=nu1, nu2=nu2, smoother=p.smoother, nthreads=4)   # mg_fmg runs V-cycles (P:646)
            assert_bitwise(mg.fullmultigrid(b, 1, nu1, nu2), orc.fullmultigrid(b, 1, pv), "fmg with zero-guess chain")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ctas", ["16", "8", "4", "1"])
@pytest.mark.parametrize("level", [3, 6, 8, 10])
def test_cluster_tail_cycles_bitwise(mgb, orc, knob, level, ctas, dtype):
    """MGB200_CTAIL=1: levels <= 8 (fewer for small clusters) in one thread-block-cluster launch (DSMEM)."""
    knob("MGB200_CTAIL", "1")
    knob("MGB200_CTAIL_CTAS", ctas)
    x, b = rand_vec(level, dtype, 85), rand_vec(level, dtype, 86, 1e-3)
    for smoother, nu1, nu2, gamma in (("jacobi", 2, 2, 1), ("jacobi", 1, 2, 2), ("rbgs", 2, 2, 1), ("rbgs", 1, 1, 2)):
        p = oracle.Params(nu1=nu1, nu2=nu2, gamma=gamma, smoother=1 if smoother == "rbgs" else 0, nthreads=4)
        want = [x]
        for _ in range(2):
            want.append(orc.vcyclemultigrid(want[-1], b, p))
        for graph in (False, True):
            with mgb.Multigrid(level, dtype=dtype, smoother=smoother, graph=graph) as mg:
                mg.set_u(level, x)
                mg.set_rhs(level, b)
                for k in range(2):
                    mg.cycle(level, nu1, nu2, gamma)
                    assert_bitwise(mg.get_u(level), want[k + 1], f"ctail cycle {k + 1} C={ctas} graph={graph} {smoother} g={gamma}")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("smoother,nu1,nu2", [("jacobi", 2, 2), ("jacobi", 1, 1), ("jacobi", 1, 2), ("jacobi", 2, 1), ("jacobi", 3, 1),
                                              ("rbgs", 1, 1), ("rbgs", 2, 2)])
@pytest.mark.parametrize("level", [3, 5, 7, 8, 10])
def test_visit_chain_postpre_bitwise(mgb, orc, knob, level, dtype, smoother, nu1, nu2):
    """MGB200_CHAIN=1: POST of one visit of a level and PRE of the next visit are one POSTPRE launch -- consecutive cycles
    through mg_cycles (the loop P:646-648) and the gamma visits of a W-cycle through mg_cycle; fullmultigrid uses it per
    level.  Same bits as the same number of separate cycles."""
    knob("MGB200_CHAIN", "1")
    x, b = rand_vec(level, dtype, 87), rand_vec(level, dtype, 88, 1e-3)
    sid = 1 if smoother == "rbgs" else 0
    for gamma, count in ((1, 3), (2, 1), (2, 2)):
        p = oracle.Params(nu1=nu1, nu2=nu2, gamma=gamma, smoother=sid, nthreads=4)
        want = [x]
        for _ in range(2 * count):
            want.append(orc.vcyclemultigrid(want[-1], b, p))
        for graph, tail in ((False, False), (True, True)):
            with mgb.Multigrid(level, dtype=dtype, smoother=smoother, graph=graph, coarse_tail=tail) as mg:
                mg.set_u(level, x)
                mg.set_rhs(level, b)
                mg.cycles(count, level, nu1, nu2, gamma)
                assert_bitwise(mg.get_u(level), want[count], f"chain g={gamma} n={count} graph={graph} tail={tail}")
                mg.cycles(count, level, nu1, nu2, gamma)      # replay from the new buffer parities
                assert_bitwise(mg.get_u(level), want[2 * count], f"chain replay g={gamma} n={count} graph={graph}")
    # fullmultigrid: the interpolation of the coarse solution (P:645) is fused into the first PRE of each level
    # (k_stream_fmg_entry), its cycles per level are chained
    pv = oracle.Params(nu1=nu1, nu2=nu2, smoother=sid, nthreads=4)
    for graph, tail in ((False, False), (True, True)):
        with mgb.Multigrid(level, dtype=dtype, smoother=smoother, graph=graph, coarse_tail=tail) as mg:
            for cyc in (1, 3):
                assert_bitwise(mg.fullmultigrid(b, cyc, nu1, nu2), orc.fullmultigrid(b, cyc, pv), f"fmg, {cyc} cycles per level")
            # a pending interpolation must be materialised for any other reader
            if level > 2:
                mg.set_rhs(level, b)
                mg.fmg(1, nu1, nu2)
                assert_bitwise(mg.get_u(level), orc.fullmultigrid(b, 1, pv), "resident fmg")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("smoother,nu1,nu2,gamma", [("jacobi", 2, 2, 1), ("jacobi", 1, 1, 2), ("jacobi", 3, 5, 1), ("rbgs", 2, 2, 1), ("rbgs", 1, 1, 2)])
@pytest.mark.parametrize("level", [3, 6, 7, 9, 10])
def test_tma_streaming_kernels_bitwise(mgb, orc, knob, level, dtype, smoother, nu1, nu2, gamma):
    """MGB200_TMA=1: the streaming kernels fetch rows with 1-D bulk copies (cp.async.bulk, UBLKCP) completing on one
    mbarrier per ring slot instead of per-lane cp.async.  Same pipeline, same bits.  (On a GPU run this first, alone,
    under a short timeout: a wrong transaction count would hang the kernel.)"""
    knob("MGB200_TMA", "1")
    x, b = rand_vec(level, dtype, 89), rand_vec(level, dtype, 90, 1e-3)
    p = oracle.Params(nu1=nu1, nu2=nu2, gamma=gamma, smoother=1 if smoother == "rbgs" else 0, nthreads=4)
    want = [x]
    for _ in range(2):
        want.append(orc.vcyclemultigrid(want[-1], b, p))
    for graph, tail in ((False, False), (True, True)):
        with mgb.Multigrid(level, dtype=dtype, smoother=smoother, graph=graph, coarse_tail=tail) as mg:
            mg.set_u(level, x)
            mg.set_rhs(level, b)
            for k in range(2):
                mg.cycle(level, nu1, nu2, gamma)
                assert_bitwise(mg.get_u(level), want[k + 1], f"tma cycle {k + 1} graph={graph} tail={tail}")
            mg.set_u(level, x)
            mg.smooth(level, 3)
            ref = orc.jacobirelaxation(x, b, 3) if smoother == "jacobi" else orc.rbgs(x, b, 3)
            assert_bitwise(mg.get_u(level), ref, "tma sweeps")


def test_visit_chain_really_fuses(mgb, knob):
    """Launch counts: 3 chained V(2,2) cycles at 513^2 save two launches on the finest level, a W-cycle one per level."""
    counts = {}
    for chain in ("0", "1"):
        knob("MGB200_CHAIN", chain)
        for gamma, n in ((1, 3), (2, 1)):
            with mgb.Multigrid(9, graph=False) as mg:
                mg.force_constant(4.0)
                mg.zero_u(9)
                l0 = mg.launches
                mg.cycles(n, 9, 2, 2, gamma)
                counts[chain, gamma] = mg.launches - l0
    assert counts["1", 1] == counts["0", 1] - 2 and counts["1", 2] < counts["0", 2]


def test_tile_kernels_full_size_and_speed(mgb, orc, knob):
    """4097^2: the tile kernels take over levels <= 10; result must not change, cycle must not get slower."""
    level = 12
    n = (1 << level) - 1
    b = (1.0 / 4096.0) ** 2 * np.random.default_rng(1234).uniform(-1, 1, n * n)
    want = orc.vcyclemultigrid(np.zeros(n * n), b, oracle.Params(nthreads=orc.max_threads()))
    times = {}
    for tile in ("0", "1"):
        knob("MGB200_TILE", tile)
        with mgb.Multigrid(level) as mg:
            mg.set_rhs(level, b)
            mg.zero_u(level)
            mg.cycle(level, 2, 2, 1)
            assert_bitwise(mg.get_u(level), want, f"V(2,2) at 4097^2, MGB200_TILE={tile}")
            mg.time_cycle(level, 2, 2, 1, 5)
            times[tile] = mg.time_cycle(level, 2, 2, 1, 20) / 20
    print(f"V(2,2) 4097^2: stream-only {times['0'] * 1e3:.1f} us, with tile kernels {times['1'] * 1e3:.1f} us")
    assert times["1"] < 1.05 * times["0"]


@pytest.mark.parametrize("env", [{"MGB200_COMM_AVOID": "1"}, {"MGB200_GRAPH_DIST": "1"},
                                 {"MGB200_COMM_AVOID": "1", "MGB200_GRAPH_DIST": "1"}, {"MGB200_TILE": "1"}])
def test_multigpu_optin_paths(env):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env={**os.environ, **env})
    assert out.returncode == 0 and "MGPU OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs 3-5 at FULL size on one GPU (minutes of oracle time and tens of GB of host memory:
# opt-in).  Bit-exact where the oracle finishes in reasonable time, size-independent properties otherwise.
# ------------------------------------------------------------------------------------------------
def test_config3_16385_rbgs_vcycle_bitwise(mgb, orc):
    level = 14
    n = (1 << level) - 1
    b = (1.0 / (1 << level)) ** 2 * np.random.default_rng(1234).uniform(-1, 1, n * n)
    p = oracle.Params(smoother=1, nthreads=orc.max_threads())
    with mgb.Multigrid(level, smoother="rbgs") as mg:
        mg.set_rhs(level, b)
        mg.zero_u(level)
        r0 = mg.residual(level, norm=True)
        mg.cycle(level, 2, 2, 1)
        r1 = mg.residual(level, norm=True)
        assert_bitwise(mg.get_u(level), orc.vcyclemultigrid(np.zeros(n * n), b, p), "RB-GS V(2,2) at 16385^2")
        assert r1 / r0 < 0.1


def test_config4_8193_wcycle_and_fmg_bitwise(mgb, orc):
    level = 13
    n = (1 << level) - 1
    b = (1.0 / (1 << level)) ** 2 * np.random.default_rng(1234).uniform(-1, 1, n * n)
    nt = orc.max_threads()
    with mgb.Multigrid(level) as mg:
        mg.set_rhs(level, b)
        mg.zero_u(level)
        mg.cycle(level, 2, 2, 2)
        assert_bitwise(mg.get_u(level), orc.vcyclemultigrid(np.zeros(n * n), b, oracle.Params(gamma=2, nthreads=nt)), "W(2,2) at 8193^2")
        assert_bitwise(mg.fullmultigrid(b, 1, 2, 2), orc.fullmultigrid(b, 1, oracle.Params(nthreads=nt)), "FMG at 8193^2")


def test_config5_32769_fp32_smoother_residual_properties(mgb):
    """32769^2 fp32 (4.3 GB per array): linearity of smoother and residual under exact scalings, and temporal
    blocking (two sweeps in one launch) equals two single sweeps."""
    level = 15
    with mgb.Multigrid(level, coarsest_level=level - 1, dtype=np.float32) as mg:
        mg.force_constant(4.0)                      # b = 4 h^2 (exact in fp32)
        mg.zero_u(level)
        mg.smooth(level, 2)
        r1 = mg.residual(level, norm=True)
        mg.force_constant(8.0)                      # scaling by 2 is exact: every iterate and the norm double
        mg.zero_u(level)
        mg.smooth(level, 2)
        r2 = mg.residual(level, norm=True)
        assert r2 == 2.0 * r1
        mg.zero_u(level)
        mg.smooth(level, 1)
        mg.smooth(level, 1)
        assert mg.residual(level, norm=True) == r2
