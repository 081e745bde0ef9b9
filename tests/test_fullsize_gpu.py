"""BASELINE.json configs 3-5 at FULL size on one GPU (parity tests proper, through the C ABI).  Bit-exact against the oracle
where the oracle finishes in seconds with all host threads (16385^2 RB-GS V-cycle, 8193^2 W-cycle and full multigrid),
size-independent properties where it does not (32769^2 fp32: exact scalings, temporal blocking == separate sweeps).
They need a few GB of host memory and ~1-2 minutes; marked `slow` so that `-m "gpu and not slow"` skips them."""
import numpy as np
import pytest

import oracle
from conftest import assert_bitwise

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


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
