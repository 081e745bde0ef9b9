"""CPU-only tests that pin the oracle (SURVEY.md Appendix C known answers; the reference
itself ships no tests or golden vectors)."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

import oracle
from conftest import rand_vec

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_c4_commented_smoke_test_exact_rationals(orc):
    """P:694-722 worked out with the reference RHS: N=9, b=4h^2, v0=0, 2 sweeps, w=2/3.
    Exact values 5/288 (corner), 11/576 (edge midpoint), 1/48 (centre)."""
    n = 7
    b = orc.globalforcefunction(3)
    assert np.all(b == 4.0 / 64.0)
    v = orc.jacobirelaxation(np.zeros(n * n), b, 2).reshape(n, n)
    for got, exact in ((v[0, 0], Fraction(5, 288)), (v[0, 3], Fraction(11, 576)), (v[3, 3], Fraction(1, 48))):
        assert abs(got - float(exact)) <= 2 * np.spacing(float(exact))
    # the as-printed values of SURVEY C.4
    assert v[0, 0] == pytest.approx(0.017361111111111112, rel=1e-15)
    assert v[3, 3] == pytest.approx(0.020833333333333332, rel=1e-15)


def test_c4_f2_variant(orc):
    """P:711-715 literally: 49 unknowns, f_h = 2, v0 = 0, 2 sweeps -> 5/9, 11/18, 2/3."""
    n = 7
    v = orc.jacobirelaxation(np.zeros(n * n), np.full(n * n, 2.0), 2).reshape(n, n)
    assert v[0, 0] == pytest.approx(5 / 9, rel=1e-15)
    assert v[0, 3] == pytest.approx(11 / 18, rel=1e-15)
    assert v[3, 3] == pytest.approx(2 / 3, rel=1e-15)


CASES_C3 = [
    # smoother, gamma, rhs, cycles, final relres, first factors, asymptotic factor
    (0, 1, "const", 13, 7.287e-09, [0.2488, 0.2380, 0.2369, 0.2359], 0.2353),
    (0, 1, "rand", 12, 2.227e-09, [0.0840, 0.1795, 0.1901, 0.1972], 0.2166),
    (1, 1, "const", 7, 4.102e-09, [0.0790, 0.0591, 0.0606, 0.0613], 0.0619),
    (0, 2, "const", 10, 3.803e-09, [0.0897, 0.0396, 0.1520, 0.1738], 0.1896),
    (1, 2, "const", 5, 1.317e-09, [0.0043, 0.0179, 0.0218, 0.0258], 0.0304),
]


@pytest.mark.parametrize("smoother,gamma,rhs,cycles,relres,first,asym", CASES_C3)
def test_c3_convergence_257(orc, smoother, gamma, rhs, cycles, relres, first, asym):
    level, n = 8, 255
    if rhs == "const":
        b = orc.globalforcefunction(level)
    else:
        b = (1.0 / 256.0) ** 2 * np.random.default_rng(1234).uniform(-1, 1, n * n)
    p = oracle.Params(smoother=smoother, gamma=gamma, nthreads=orc.max_threads())
    u, k, hist = orc.solve(np.zeros(n * n), b, 1e-8, 60, p)
    fac = hist[1:] / hist[:-1]
    assert k == cycles
    assert hist[-1] / hist[0] == pytest.approx(relres, rel=2e-3)
    assert np.allclose(fac[:4], first, atol=6e-5)
    assert fac[-1] == pytest.approx(asym, abs=6e-5)
    if smoother == 0 and gamma == 1 and rhs == "const":
        assert u.max() == pytest.approx(0.294681869172, rel=1e-11)
        assert u.reshape(n, n)[127, 127] == pytest.approx(0.29468541, rel=2e-5)  # analytic u(1/2,1/2)


def test_c3_1025_and_fmg(orc):
    nt = orc.max_threads()
    b = orc.globalforcefunction(10)
    u, k, hist = orc.solve(np.zeros(b.size), b, 1e-8, 60, oracle.Params(nthreads=nt))
    assert k == 13 and hist[-1] / hist[0] == pytest.approx(8.876e-09, rel=2e-3)
    assert (hist[1:] / hist[:-1])[-1] == pytest.approx(0.2385, abs=6e-5)
    b8 = orc.globalforcefunction(8)
    ufmg = orc.fullmultigrid(b8, 1, oracle.Params(nthreads=nt))
    assert orc.norm2(orc.residual(ufmg, b8)) / orc.norm2(b8) == pytest.approx(1.782e-02, rel=2e-3)


def test_reference_depth_and_literal_scaling(orc):
    """E4/E5 of SURVEY App. A: the literal 1/16 weight on unscaled stencils stalls (~0.99 per
    cycle); the reference depth (nu=10, L_min=L-3) with w=1/4 converges at ~0.8-0.92."""
    nt = orc.max_threads()
    b = orc.globalforcefunction(8)
    _, _, h = orc.solve(np.zeros(b.size), b, 1e-8, 8, oracle.Params(restrict_weight=1.0 / 16.0, nthreads=nt))
    f = h[1:] / h[:-1]
    assert f[0] == pytest.approx(0.9865, abs=2e-4) and f[1] == pytest.approx(0.9910, abs=2e-4)
    _, _, h = orc.solve(np.zeros(b.size), b, 1e-8, 8, oracle.Params(nu1=10, nu2=10, coarsest_level=5, nthreads=nt))
    f = h[1:] / h[:-1]
    assert f[0] == pytest.approx(0.794, abs=2e-3) and f[1] == pytest.approx(0.890, abs=2e-3)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_c2_transfer_identities(orc, dtype):
    """interpolation2d is bilinear with a zero ring, restriction2d(1/16) is full weighting,
    and <P x, y> = 4 <x, R_{1/16} y> (P = 4 R^T)."""
    cl, fl = 4, 5
    m, n = 15, 31
    # data with 16-bit mantissas so that every sum below is exact in fp32 too
    x = (np.random.default_rng(7).integers(-2 ** 12, 2 ** 12, m * m) / 2.0 ** 8).astype(dtype)
    y = (np.random.default_rng(8).integers(-2 ** 12, 2 ** 12, n * n) / 2.0 ** 8).astype(dtype)
    px = orc.interpolation2d(x)
    xp = np.zeros((m + 2, m + 2), dtype=dtype)
    xp[1:-1, 1:-1] = x.reshape(m, m)
    ref = np.zeros((n + 2, n + 2), dtype=dtype)
    ref[0::2, 0::2] = xp
    ref[1::2, 0::2] = 0.5 * (xp[:-1, :] + xp[1:, :])
    ref[0::2, 1::2] = 0.5 * (xp[:, :-1] + xp[:, 1:])
    ref[1::2, 1::2] = 0.25 * (xp[:-1, :-1] + xp[1:, :-1] + xp[:-1, 1:] + xp[1:, 1:])
    assert np.array_equal(px.reshape(n, n), ref[1:-1, 1:-1])
    ry = orc.restriction2d(y, w=1.0 / 16.0)
    Y = y.reshape(n, n).astype(np.float64)
    fw = (Y[0:-2:2, 0:-2:2] + Y[0:-2:2, 2::2] + Y[2::2, 0:-2:2] + Y[2::2, 2::2]
          + 2 * (Y[1:-1:2, 0:-2:2] + Y[1:-1:2, 2::2] + Y[0:-2:2, 1:-1:2] + Y[2::2, 1:-1:2])
          + 4 * Y[1:-1:2, 1:-1:2]) / 16.0
    assert np.array_equal(ry.reshape(m, m).astype(np.float64), fw)
    lhs = float(np.dot(px.astype(np.float64), y.astype(np.float64)))
    rhs = float(np.dot(x.astype(np.float64), ry.astype(np.float64)))
    assert lhs / rhs == pytest.approx(4.0, rel=1e-12)
    # constants are preserved by R (weight 1/16) away from the boundary
    ones = orc.restriction2d(np.ones(n * n, dtype=dtype), w=1.0 / 16.0)
    assert np.all(ones == 1.0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_residual_of_converged_solution_and_correction(orc, dtype):
    level = 5
    b = orc.globalforcefunction(level, dtype=dtype)
    u, k, h = orc.solve(np.zeros_like(b), b, 2e-5 if dtype == np.float32 else 1e-12, 60)
    assert h[-1] / h[0] < (2e-5 if dtype == np.float32 else 1e-12)
    e = rand_vec(level - 1, dtype, 3)
    assert np.array_equal(orc.prolong_correct(e, u), u + orc.interpolation2d(e))


def test_csr_reference_structured_path_matches_matrix_free(orc):
    """CPU baseline A (CSR SpMV + scal/add passes as P:138-144, P:604-607) computes the same
    iterates as the matrix-free oracle up to the SpMV row-sum order."""
    level = 6
    hd = {l: orc.csr_build(l) for l in range(1, level + 1)}
    x, b = rand_vec(level, np.float64, 1), rand_vec(level, np.float64, 2, 1e-3)
    assert np.allclose(orc.csr_jacobirelaxation(hd[level], x, b, 3), orc.jacobirelaxation(x, b, 3), rtol=0, atol=1e-15)
    assert np.allclose(orc.csr_residual(hd[level], x, b), orc.residual(x, b), rtol=0, atol=1e-14)
    assert np.allclose(orc.csr_vcyclemultigrid(hd, x, b), orc.vcyclemultigrid(x, b), rtol=0, atol=1e-13)
    for h in hd.values():
        orc.csr_free(h)


def test_fp32_constants_follow_reference_types(orc):
    """P:127 `const float omega = 2.0/3.0`; P:138-140 form 1-omega and omega/4 in double from
    the float omega and the oneMKL call narrows them to float."""
    om = np.float32(2.0 / 3.0)
    c0, c1 = orc.jacobi_constants(2.0 / 3.0, np.float32)
    assert np.float32(c0) == np.float32(1.0 - float(om)) and np.float32(c1) == np.float32(float(om) / 4.0)
    c0d, c1d = orc.jacobi_constants(2.0 / 3.0, np.float64)
    assert c0d == 1.0 - 2.0 / 3.0 and c1d == (2.0 / 3.0) / 4.0


def test_threads_do_not_change_results(orc):
    level = 7
    x, b = rand_vec(level, np.float64, 11), rand_vec(level, np.float64, 12, 1e-4)
    p1, p8 = oracle.Params(nthreads=1), oracle.Params(nthreads=max(2, orc.max_threads()))
    assert np.array_equal(orc.vcyclemultigrid(x, b, p1), orc.vcyclemultigrid(x, b, p8))
    p1.smoother = p8.smoother = 1
    assert np.array_equal(orc.vcyclemultigrid(x, b, p1), orc.vcyclemultigrid(x, b, p8))


def test_golden_fixtures_if_present(orc):
    """tests/golden/*.npz hold oracle outputs frozen at commit time (generated by
    tests/golden/make_golden.py); they catch accidental changes of the oracle itself."""
    path = os.path.join(GOLD, "oracle_golden.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixtures not generated")
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    for case in meta["cases"]:
        name, level, dtype = case["name"], case["level"], np.dtype(case["dtype"])
        x = rand_vec(level, dtype, case["seed_u"])
        b = rand_vec(level, dtype, case["seed_b"], case["scale_b"])
        p = oracle.Params(smoother=case["smoother"], gamma=case["gamma"])
        got = {"jacobi3": lambda: orc.jacobirelaxation(x, b, 3),
               "rbgs2": lambda: orc.rbgs(x, b, 2),
               "residual": lambda: orc.residual(x, b),
               "restrict": lambda: orc.restriction2d(x),
               "prolong": lambda: orc.interpolation2d(x),
               "vcycle": lambda: orc.vcyclemultigrid(x, b, p),
               "fmg": lambda: orc.fullmultigrid(b, 1, p)}[case["op"]]()
        assert np.array_equal(got, g[name]), name


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-3)])
@pytest.mark.parametrize("level", [1, 2, 4, 6])
def test_exact_coarse_solve_solves_the_five_point_system(orc, level, dtype, tol):
    """mgo_coarse_exact (direct_solver, M:63-72): A u = f to rounding, and A^-1 of the constant load is symmetric."""
    n = (1 << level) - 1
    f = np.random.default_rng(level).uniform(-1, 1, n * n).astype(dtype)
    u = orc.coarse_exact(f)
    assert np.abs(orc.residual(u, f)).max() <= tol * max(1.0, float(np.abs(f).max()))
    c = orc.coarse_exact(np.ones(n * n, dtype=dtype)).reshape(n, n)
    assert np.allclose(c, c.T, rtol=0, atol=tol) and np.allclose(c, c[::-1, ::-1], rtol=0, atol=tol)
    if level == 1:
        assert orc.coarse_exact(np.array([1.0], dtype=dtype))[0] == pytest.approx(0.25, rel=1e-6)


def test_exact_coarse_solve_fixes_the_reference_depth(orc):
    """coarsest = finest - 3 (P:17-18): sweeps on the coarsest grid stall (factor ~0.97), the exact solve of the
    second version (M:136-139) gives the textbook V(2,2) factor, the same as coarsening down to one unknown."""
    import oracle
    level = 8
    n = (1 << level) - 1
    b = np.full(n * n, 4.0 / (1 << level) ** 2)
    fac = {}
    for name, p in (("sweeps", oracle.Params(coarsest_level=5, nthreads=4)),
                    ("exact", oracle.Params(coarsest_level=5, nthreads=4, coarse_exact=1)),
                    ("deep", oracle.Params(coarsest_level=1, nthreads=4))):
        u = np.zeros(n * n)
        h = [orc.norm2(orc.residual(u, b))]
        for _ in range(8):
            u = orc.vcyclemultigrid(u, b, p)
            h.append(orc.norm2(orc.residual(u, b)))
        fac[name] = h[-1] / h[-2]
    assert fac["sweeps"] > 0.9 and 0.18 < fac["exact"] < 0.25 and abs(fac["exact"] - fac["deep"]) < 0.03, fac
