"""GPU parity tests: every CUDA entry point of libmgb200.so (called through the C ABI) is
compared with the CPU oracle on identical seeded inputs.  The library is built with
--fmad=false and evaluates the point formulas in the oracle's order, so the bar is
BIT-EXACT for every operator and whole cycles (tighter than the north star's 1e-12
relative); only the residual norm (a reduction) carries a tolerance, written below."""
import json
import os

import numpy as np
import pytest

import conftest
import oracle
from conftest import assert_bitwise, rand_vec

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32]
NORM_RTOL = 1e-13   # GPU tree reduction vs oracle's sequential double sum


def make(mgb, level, dtype=np.float64, **kw):
    kw.setdefault("coarsest_level", 1)
    return mgb.Multigrid(level, dtype=dtype, **kw)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 7, 8, 9, 10])
def test_jacobi_iterates_bitwise(mgb, orc, level, dtype):
    x, b = rand_vec(level, dtype, 1), rand_vec(level, dtype, 2, 1e-3)
    with make(mgb, level, dtype, coarsest_level=min(level, 1)) as mg:
        for nu in (1, 2, 3, 5):
            mg.set_u(level, x)
            mg.set_rhs(level, b)
            mg.smooth(level, nu)
            assert_bitwise(mg.get_u(level), orc.jacobirelaxation(x, b, nu), f"jacobi L{level} nu={nu}")
        # the reference-shaped host call (P:125): mutates v and returns it
        v = x.copy()
        out = mg.jacobirelaxation(v, b, 4)
        assert_bitwise(out, orc.jacobirelaxation(x, b, 4), "host jacobirelaxation")
        assert_bitwise(v, out, "v mutated in place (P:146)")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("level", [1, 2, 3, 5, 8, 10])
def test_rbgs_bitwise(mgb, orc, level, dtype):
    x, b = rand_vec(level, dtype, 3), rand_vec(level, dtype, 4, 1e-3)
    with make(mgb, level, dtype, smoother="rbgs") as mg:
        for nu in (1, 2, 3):
            mg.set_u(level, x)
            mg.set_rhs(level, b)
            mg.smooth(level, nu)
            assert_bitwise(mg.get_u(level), orc.rbgs(x, b, nu), f"rbgs L{level} nu={nu}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("level", [1, 2, 3, 6, 9, 10])
def test_residual_and_norm(mgb, orc, level, dtype):
    x, b = rand_vec(level, dtype, 5), rand_vec(level, dtype, 6)
    with make(mgb, level, dtype) as mg:
        mg.set_u(level, x)
        mg.set_rhs(level, b)
        nrm = mg.residual(level, norm=True)
        r = orc.residual(x, b)
        assert_bitwise(mg.get_r(level), r, f"residual L{level}")
        assert nrm == pytest.approx(orc.norm2(r), rel=NORM_RTOL if dtype == np.float64 else 1e-12)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("level", [2, 3, 4, 6, 9, 10])
def test_transfer_operators_bitwise(mgb, orc, level, dtype):
    fine, coarse = rand_vec(level, dtype, 7), rand_vec(level - 1, dtype, 8)
    with make(mgb, level, dtype) as mg:
        assert_bitwise(mg.restriction2d(fine), orc.restriction2d(fine), "restriction2d")
        assert_bitwise(mg.interpolation2d(coarse), orc.interpolation2d(coarse), "interpolation2d")
        # resident: prolong + correct (P:620-624)
        mg.set_u(level, fine)
        mg.set_u(level - 1, coarse)
        mg.prolong_correct(level)
        assert_bitwise(mg.get_u(level), orc.prolong_correct(coarse, fine), "prolong_correct")
        # restriction of the RHS for FMG (P:641) and the zero coarse guess (P:613)
        mg.set_rhs(level, fine)
        mg.restrict_rhs(level)
        assert_bitwise(mg.get_rhs(level - 1), orc.restriction2d(fine), "restrict_rhs")
        mg.set_u(level, fine)
        mg.set_rhs(level, rand_vec(level, dtype, 9))
        mg.residual(level)
        mg.restrict(level)
        assert_bitwise(mg.get_rhs(level - 1), orc.restriction2d(orc.residual(fine, rand_vec(level, dtype, 9))), "restrict(residual)")
        assert not mg.get_u(level - 1).any()


def test_literal_fd_weight(mgb, orc):
    level = 6
    fine = rand_vec(level, np.float64, 10)
    with make(mgb, level, restrict_weight=1.0 / 16.0) as mg:
        assert_bitwise(mg.restriction2d(fine), orc.restriction2d(fine, w=1.0 / 16.0), "restriction2d 1/16")


CYCLES = [  # smoother, nu1, nu2, gamma, coarsest
    ("jacobi", 2, 2, 1, 1), ("jacobi", 1, 1, 1, 1), ("jacobi", 2, 1, 1, 1), ("jacobi", 3, 0, 1, 2),
    ("jacobi", 2, 2, 2, 1), ("rbgs", 2, 2, 1, 1), ("rbgs", 1, 1, 2, 1), ("jacobi", 10, 10, 1, 4),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("smoother,nu1,nu2,gamma,coarsest", CYCLES)
@pytest.mark.parametrize("level", [4, 7, 9])
def test_cycles_bitwise_all_flag_combinations(mgb, orc, level, dtype, smoother, nu1, nu2, gamma, coarsest):
    """vcyclemultigrid P:575-627 (gamma=2: W): three consecutive cycles, with and without
    CUDA graphs / fused kernels / the coarse tail -- all must equal the oracle bit for bit."""
    x, b = rand_vec(level, dtype, 21), rand_vec(level, dtype, 22, 1e-3)
    p = oracle.Params(coarsest_level=coarsest, nu1=nu1, nu2=nu2, gamma=gamma, smoother=1 if smoother == "rbgs" else 0,
                      nthreads=4)
    want = [x]
    for _ in range(3):
        want.append(orc.vcyclemultigrid(want[-1], b, p))
    for graph, fused, tail in ((False, False, False), (True, False, False), (False, True, False),
                               (False, False, True), (True, True, True)):
        with make(mgb, level, dtype, coarsest_level=coarsest, smoother=smoother, graph=graph, fused=fused,
                  coarse_tail=tail) as mg:
            mg.set_u(level, x)
            mg.set_rhs(level, b)
            for k in range(3):
                mg.cycle(level, nu1, nu2, gamma)
                assert_bitwise(mg.get_u(level), want[k + 1], f"cycle {k + 1} graph={graph} fused={fused} tail={tail}")
            # host-vector entry point (one call == one reference call)
            assert_bitwise(mg.vcyclemultigrid(x, b, nu1, nu2, gamma), want[1], "host vcyclemultigrid")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("smoother", ["jacobi", "rbgs"])
def test_fullmultigrid_bitwise(mgb, orc, dtype, smoother):
    level = 8
    b = rand_vec(level, dtype, 31, 1e-3)
    for cycles, nu, coarsest in ((1, 2, 1), (2, 2, 1), (3, 10, 5)):
        p = oracle.Params(coarsest_level=coarsest, nu1=nu, nu2=nu, smoother=1 if smoother == "rbgs" else 0, nthreads=4)
        with make(mgb, level, dtype, coarsest_level=coarsest, smoother=smoother) as mg:
            assert_bitwise(mg.fullmultigrid(b, cycles, nu, nu), orc.fullmultigrid(b, cycles, p), f"fmg c={cycles}")


@pytest.mark.parametrize("smoother,gamma,cycles,relres", [("jacobi", 1, 13, 7.287e-09), ("rbgs", 1, 7, 4.102e-09),
                                                           ("jacobi", 2, 10, 3.803e-09), ("rbgs", 2, 5, 1.317e-09)])
def test_config0_solve_257_matches_reference_history(mgb, orc, smoother, gamma, cycles, relres):
    """BASELINE.json configs[0]: 257^2 fp64 V(2,2) to 1e-8: same cycle count,
    same residual history as the oracle to 1e-10 relative, same iterate bit for bit."""
    level = 8
    with make(mgb, level, smoother=smoother) as mg:
        b = mg.globalforcefunction(4.0)
        assert_bitwise(b, orc.globalforcefunction(level), "globalforcefunction")
        mg.zero_u(level)
        k, rel, hist = mg.solve(1e-8, 60, 2, 2, gamma)
        p = oracle.Params(smoother=1 if smoother == "rbgs" else 0, gamma=gamma, nthreads=4)
        u, ko, ho = orc.solve(np.zeros_like(b), b, 1e-8, 60, p)
        assert k == ko == cycles
        assert rel == pytest.approx(relres, rel=2e-3)
        assert np.allclose(hist, ho, rtol=1e-10, atol=0)
        assert_bitwise(mg.get_u(level), u, "solution")


@pytest.mark.parametrize("smoother,gamma", [("jacobi", 1), ("rbgs", 2)])
def test_mg_cycles_equals_repeated_cycles(mgb, orc, smoother, gamma):
    """mg_cycles(count) == the loop P:646-648 of `count` vcyclemultigrid calls, bit for bit (one graph for the run)."""
    level = 7
    x, b = rand_vec(level, np.float64, 51), rand_vec(level, np.float64, 52, 1e-3)
    p = oracle.Params(smoother=1 if smoother == "rbgs" else 0, gamma=gamma, nthreads=4)
    want = x
    for _ in range(4):
        want = orc.vcyclemultigrid(want, b, p)
    for graph in (False, True):
        with make(mgb, level, smoother=smoother, graph=graph) as mg:
            mg.set_u(level, x)
            mg.set_rhs(level, b)
            mg.cycles(3, level, 2, 2, gamma)
            mg.cycles(1, level, 2, 2, gamma)
            mg.cycles(0, level, 2, 2, gamma)
            assert_bitwise(mg.get_u(level), want, f"mg_cycles graph={graph}")
            with pytest.raises(mgb.capi.MgError):
                mg.cycles(-1, level)


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
@pytest.mark.parametrize("smoother,nu1,nu2,gamma", [("jacobi", 2, 2, 1), ("jacobi", 1, 2, 2), ("rbgs", 2, 2, 1), ("rbgs", 1, 1, 2),
                                                    ("jacobi", 3, 2, 1)])
@pytest.mark.parametrize("level", [4, 7, 8, 10])
@pytest.mark.parametrize("zg", ["1", "0"])
def test_zero_guess_chain_bitwise(mgb, orc, knob, level, dtype, smoother, nu1, nu2, gamma, zg):
    """Zero-guess chain (default; MGB200_ZERO_GUESS=0 is the A/B switch): PRE skips the zero coarse guess store (P:613), the
    next level's PRE / the tail do not read u.  Both settings give the oracle's bits."""
    knob("MGB200_ZERO_GUESS", zg)
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
@pytest.mark.parametrize("smoother,nu1,nu2", [("jacobi", 2, 2), ("jacobi", 1, 1), ("jacobi", 1, 2), ("jacobi", 2, 1), ("jacobi", 3, 1),
                                              ("rbgs", 1, 1), ("rbgs", 2, 2)])
@pytest.mark.parametrize("level", [3, 5, 7, 8, 10])
@pytest.mark.parametrize("chain", ["1", "0"])
def test_visit_chain_postpre_bitwise(mgb, orc, knob, level, dtype, smoother, nu1, nu2, chain):
    """Visit chains (default; MGB200_CHAIN=0 is the A/B switch): POST of one visit of a level and PRE of the next visit are one POSTPRE launch -- consecutive cycles
    through mg_cycles (the loop P:646-648) and the gamma visits of a W-cycle through mg_cycle; fullmultigrid uses it per
    level.  Same bits as the same number of separate cycles."""
    knob("MGB200_CHAIN", chain)
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


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("level", [1, 3, 6, 9])
def test_synthetic_rhs_and_checksum_match_their_numpy_restatements(mgb, level, dtype):
    """mg_force_synthetic / mg_checksum (bench.py's device-side right-hand side and its multi-GPU parity record) against
    tests/synth_ref.py, bit for bit; the checksum is sensitive to a single flipped bit and to a swap of two values."""
    import synth_ref
    with make(mgb, level, dtype) as mg:
        mg.force_synthetic(1234)
        f = mg.get_rhs(level)
        assert_bitwise(f, synth_ref.synthetic_rhs(level, 1234, dtype), "synthetic rhs")
        assert mg.checksum(level, 1) == synth_ref.checksum(level, f)
        x = rand_vec(level, dtype, 7)
        mg.set_u(level, x)
        assert mg.checksum(level, 0) == synth_ref.checksum(level, x)
        if x.size >= 2:
            y = x.copy()
            y[0], y[1] = x[1], x[0]
            mg.set_u(level, y)
            assert mg.checksum(level, 0) == synth_ref.checksum(level, y) != synth_ref.checksum(level, x)
        y = x.copy()
        y.view(np.uint64 if dtype == np.float64 else np.uint32)[x.size // 2] ^= 1
        mg.set_u(level, y)
        assert mg.checksum(level, 0) != synth_ref.checksum(level, x)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("smoother,gamma", [("jacobi", 1), ("rbgs", 1), ("jacobi", 2)])
@pytest.mark.parametrize("level,coarsest", [(3, 3), (5, 2), (7, 4), (8, 5), (9, 6)])
def test_exact_coarsest_solve_cycles_bitwise(mgb, orc, level, coarsest, dtype, smoother, gamma):
    """MG_COARSE_EXACT (direct_solver, M:63-72, called at M:136-139): the coarsest level is solved directly (sine-transform
    diagonalisation, csrc/coarse.cuh) instead of nu1+nu2 sweeps (P:583-587).  Cycles, full multigrid and the host call on
    the coarsest level alone equal the oracle bit for bit, with and without CUDA graphs."""
    x, b = rand_vec(level, dtype, 91), rand_vec(level, dtype, 92, 1e-3)
    p = oracle.Params(coarsest_level=coarsest, gamma=gamma, smoother=1 if smoother == "rbgs" else 0, nthreads=4, coarse_exact=1)
    want = [x]
    for _ in range(2):
        want.append(orc.vcyclemultigrid(want[-1], b, p))
    for graph in (False, True):
        with mgb.Multigrid(level, coarsest_level=coarsest, dtype=dtype, smoother=smoother, graph=graph, coarse_solver="exact") as mg:
            mg.set_u(level, x)
            mg.set_rhs(level, b)
            for k in range(2):
                mg.cycle(level, 2, 2, gamma)
                assert_bitwise(mg.get_u(level), want[k + 1], f"exact-coarse cycle {k + 1} graph={graph}")
            pv = oracle.Params(coarsest_level=coarsest, smoother=p.smoother, nthreads=4, coarse_exact=1)
            assert_bitwise(mg.fullmultigrid(b, 1, 2, 2), orc.fullmultigrid(b, 1, pv), "fmg with exact coarse solve")
            bc = rand_vec(coarsest, dtype, 93)
            got = mg.vcyclemultigrid(np.zeros_like(bc), bc, 2, 2, 1)          # a cycle ON the coarsest level = the solve
            assert_bitwise(got, orc.coarse_exact(bc), "direct solve on the coarsest level")


def test_exact_coarsest_solve_restores_the_textbook_rate_at_the_reference_depth(mgb, orc):
    """The reference coarsens only three levels (coarsest = finest - 3, P:17-18).  With sweeps on that coarsest grid the
    V(2,2) factor is ~0.9-0.98; with the exact solve of its second version it is the textbook ~0.22."""
    level, coarsest = 9, 6
    hist = {}
    for cs in ("sweeps", "exact"):
        with mgb.Multigrid(level, coarsest_level=coarsest, coarse_solver=cs) as mg:
            mg.force_constant(4.0)
            mg.zero_u(level)
            k, rel, h = mg.solve(1e-8, 12)
            hist[cs] = h
    f_sweeps = hist["sweeps"][-1] / hist["sweeps"][-2]
    f_exact = hist["exact"][-1] / hist["exact"][-2]
    assert f_sweeps > 0.85 and 0.15 < f_exact < 0.26, (f_sweeps, f_exact)
    with mgb.Multigrid(level, coarsest_level=1) as mg:      # coarsening all the way down gives the same rate
        mg.force_constant(4.0)
        mg.zero_u(level)
        _, _, h1 = mg.solve(1e-8, 12)
    assert abs(h1[-1] / h1[-2] - f_exact) < 0.05


@pytest.mark.parametrize("level,smoother,gamma", [(8, "jacobi", 1), (8, "rbgs", 1), (9, "jacobi", 2), (10, "jacobi", 1), (5, "jacobi", 1)])
def test_solve_as_one_device_side_loop(mgb, orc, knob, level, smoother, gamma):
    """Default on one GPU (MGB200_SOLVE_GRAPH=0 selects the host loop): the tolerance loop is ONE graph launch (conditional WHILE node, the decision taken by a kernel).
    Same cycle count, same history, same iterate as the host loop and the oracle; a second solve reuses the graph; hitting
    max_cycles stops the loop too.  (On a level that has no fused POST the call silently uses the host loop.)"""
    p = oracle.Params(smoother=1 if smoother == "rbgs" else 0, gamma=gamma, nthreads=4)
    knob("MGB200_SOLVE_GRAPH", "1")
    with make(mgb, level, smoother=smoother, graph=False) as mg0:      # without MG_GRAPH: the host loop
        mg0.force_constant(4.0)
        mg0.zero_u(level)
        k0, _, h0 = mg0.solve(1e-8, 60, 2, 2, gamma)
        u0 = mg0.get_u(level)
    with make(mgb, level, smoother=smoother) as mg:
        b = mg.globalforcefunction(4.0)
        for max_cycles in (60, 3, 60):
            mg.zero_u(level)
            g0 = mg.info(mgb.capi.MG_INFO_GRAPH_LAUNCHES)
            k, rel, hist = mg.solve(1e-8, max_cycles, 2, 2, gamma)
            u, ko, ho = orc.solve(np.zeros_like(b), b, 1e-8, max_cycles, p)
            assert k == ko and np.allclose(hist, ho, rtol=1e-10, atol=0)
            assert_bitwise(mg.get_u(level), u, f"solution, max_cycles={max_cycles}")
            if max_cycles == 60:
                assert k == k0 and np.allclose(hist, h0, rtol=1e-12, atol=0)
                assert_bitwise(mg.get_u(level), u0, "device loop vs host loop")
            if level >= 8:
                assert mg.info(mgb.capi.MG_INFO_GRAPH_LAUNCHES) - g0 == 1, "the whole solve must be one graph launch"
        mg.cycle(level, 2, 2, gamma)                     # ordinary calls still work on the state the loop left
        assert_bitwise(mg.get_u(level), orc.vcyclemultigrid(u, b, p), "cycle after a device-loop solve")


def test_golden_fixtures(mgb):
    """Committed oracle outputs (tests/golden/oracle_golden.npz, made by make_golden.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz"))
    meta = json.loads(str(g["meta"]))
    for case in meta["cases"]:
        level, dtype = case["level"], np.dtype(case["dtype"])
        x = rand_vec(level, dtype, case["seed_u"])
        b = rand_vec(level, dtype, case["seed_b"], case["scale_b"])
        smoother = "rbgs" if (case["smoother"] == 1 or case["op"] == "rbgs2") else "jacobi"
        with make(mgb, level + (1 if case["op"] == "prolong" else 0), dtype, smoother=smoother) as mg:
            op = case["op"]
            if op in ("jacobi3", "rbgs2"):
                mg.set_u(level, x); mg.set_rhs(level, b); mg.smooth(level, 3 if op == "jacobi3" else 2)
                got = mg.get_u(level)
            elif op == "residual":
                mg.set_u(level, x); mg.set_rhs(level, b); mg.residual(level)
                got = mg.get_r(level)
            elif op == "restrict":
                got = mg.restriction2d(x)
            elif op == "prolong":
                got = mg.interpolation2d(x)
            elif op == "vcycle":
                got = mg.vcyclemultigrid(x, b, 2, 2, case["gamma"])
            else:
                got = mg.fullmultigrid(b, 1, 2, 2)
            assert_bitwise(got, g[case["name"]], case["name"])


def test_config1_full_size_4097_vcycle(mgb, orc):
    """BASELINE.json configs[1] at full size: one V(2,2) on 4097^2 fp64 equals the oracle
    bit for bit; plus size-independent properties (linearity of the cycle in (u, b), and
    residual reduction factor of C.3)."""
    level = 12
    n = (1 << level) - 1
    b = (1.0 / 4096.0) ** 2 * np.random.default_rng(1234).uniform(-1, 1, n * n)
    p = oracle.Params(nthreads=orc.max_threads())
    with make(mgb, level) as mg:
        mg.set_rhs(level, b)
        mg.zero_u(level)
        r0 = mg.residual(level, norm=True)
        mg.cycle(level, 2, 2, 1)
        r1 = mg.residual(level, norm=True)
        u1 = mg.get_u(level)
        assert_bitwise(u1, orc.vcyclemultigrid(np.zeros(n * n), b, p), "V(2,2) at 4097^2")
        assert 0.05 < r1 / r0 < 0.12          # first factor for a random RHS (C.3: 0.084 at 257^2)
        # linearity: cycle(0, 2b) == 2 * cycle(0, b) exactly (scaling by 2 is exact)
        mg.set_rhs(level, 2.0 * b)
        mg.zero_u(level)
        mg.cycle(level, 2, 2, 1)
        assert_bitwise(mg.get_u(level), 2.0 * u1, "linearity in b")


def test_errors_are_loud(mgb):
    with make(mgb, 5) as mg:
        with pytest.raises(mgb.capi.MgError):
            mg.smooth(9, 1)
        with pytest.raises(mgb.capi.MgError):
            mg.restrict(1)
        with pytest.raises(ValueError):
            mg.set_u(5, np.zeros(10))
    with pytest.raises(mgb.capi.MgError):
        mgb.Multigrid(3, coarsest_level=4)
    with pytest.raises(mgb.capi.MgError):                       # the sine-transform table of the exact solve is n x n: n <= 511
        mgb.Multigrid(11, coarsest_level=10, coarse_solver="exact")
    with make(mgb, 4) as mg:
        import ctypes
        out = ctypes.c_uint64(0)
        L = mgb.capi.lib()
        assert L.mg_checksum(mg._ctx, 4, 7, ctypes.byref(out)) == mgb.capi.MG_ERR_ARG      # which not in {u, f, r}
        assert L.mg_checksum(mg._ctx, 4, 0, None) == mgb.capi.MG_ERR_ARG
        assert L.mg_checksum(mg._ctx, 9, 0, ctypes.byref(out)) == mgb.capi.MG_ERR_ARG      # level outside the hierarchy
        assert b"level" in L.mg_last_error(mg._ctx)
        k, rel, hist = mg.solve(1e-8, 0)                        # max_cycles = 0: no cycle, history = the initial norm only
        assert k == 0 and len(hist) == 1
    # a failed mg_create leaves nothing behind: the next context on the same device works
    with make(mgb, 5) as mg:
        mg.force_constant(4.0)
        mg.zero_u(5)
        assert mg.solve(1e-8, 30)[0] > 0


def test_cpp_driver_example_runs_like_the_reference_main(mgb, tmp_path):
    """include/mgb200_driver.hpp: the reference's function names over the C ABI
    (examples/poisson_main.cpp == main() P:658-731).  257^2, V(2,2): 13 cycles (C.3)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "multigrid_nikhil_c-_b200", "lib")
    exe = str(tmp_path / "poisson_main")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "poisson_main.cpp"), "-o", exe, "-L" + libdir, "-lmgb200",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe, "8", "1", "2"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "Size of finest level solution is 65025" in out.stdout
    assert "V(2,2) solve: 13 cycles" in out.stdout and "Program Running Correctly" in out.stdout
    # FMG with one V(2,2) per level (mu0 = 0): the discrete solution is within 2% in the max norm
    umax = float(out.stdout.split("max u = ")[1].split()[0])
    assert abs(umax - 0.2946818) < 0.01


def test_cpp_problemvar_example(mgb, tmp_path):
    """examples/problemvar_main.cpp: the v2 call shape multigrid_solver(ProblemVar&) (M:193-197) with a sampled
    right-hand side and Dirichlet data; exact solution x^2 + y^2."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "multigrid_nikhil_c-_b200", "lib")
    exe = str(tmp_path / "problemvar_main")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "problemvar_main.cpp"), "-o", exe, "-L" + libdir, "-lmgb200",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe, "8"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "max |u - (x^2+y^2)|" in out.stdout, out.stdout + out.stderr
