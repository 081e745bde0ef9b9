"""Pins of the oracle against the REFERENCE'S OWN CODE (CPU only).

`oracle/_ref/libref_poisson.so` is /root/reference/Poissons_SYCL.cpp compiled where it lies
against stub oneMKL/SYCL headers (oracle/stub, oracle/ref_shim.cpp; `make -C oracle ref`).
Its outputs are frozen in tests/golden/ref_pins.json + ref_interp.npz by
tests/golden/make_ref_golden.py, so these tests also run where /root/reference is absent.
When the library is present the live outputs are checked against the fixtures too."""
import ctypes
import json
import os

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PINS = json.load(open(os.path.join(HERE, "golden", "ref_pins.json")))
INTERP = np.load(os.path.join(HERE, "golden", "ref_interp.npz"))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_poisson.so")


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    R = ctypes.CDLL(REF_SO)
    R.ref_globalforcefunction.restype = ctypes.c_longlong
    return R


def test_interpolation2d_matches_reference_bit_for_bit(orc):
    """P:337-425 incl. its corner statements and four boundary loops == the oracle's single
    zero-ring formula, for every size in the fixture (fp32, the reference's type)."""
    for m in (1, 2, 3, 7, 15, 31):
        x, want = INTERP[f"interp_in_{m}"], INTERP[f"interp_out_{m}"]
        assert np.array_equal(orc.interpolation2d(x), want), m


def test_interpolation2d_live_reference(orc, ref):
    for m in (1, 5, 63, 127):
        x = np.random.default_rng(m).uniform(-1, 1, m * m).astype(np.float32)
        out = np.zeros((2 * m + 1) ** 2, np.float32)
        ref.ref_interpolation2d(P(x), m, P(out))
        assert np.array_equal(out, orc.interpolation2d(x)), m


def _intended_a_lu(m):
    rp, ci = [0], []
    for r in range(m):
        for c in range(m):
            for rr, cc in ((r - 1, c), (r, c - 1), (r, c + 1), (r + 1, c)):
                if 0 <= rr < m and 0 <= cc < m:
                    ci.append(rr * m + cc)
            rp.append(len(ci))
    return np.array(rp, np.int32), np.array(ci, np.int32), np.full(len(ci), -1.0, np.float32)


def test_jacobirelaxation_matches_the_references_own_body_on_the_intended_operator(orc):
    """P:125-147 (gemv alpha = -omega/4, scal (1-omega), scal omega/4, two adds; omega = 2/3 P:127) executed by the
    reference's own function on A_lu = -1 per neighbour.  The oracle sums the four neighbours as (N+S)+(W+E), the CSR
    row sum runs N, W, E, S: same algebra, different association => a few fp32 ulps per sweep, not bits."""
    for m, mu in ((7, 1), (15, 3), (31, 10)):
        v, fh, want = INTERP[f"jacobi_v_{m}"], INTERP[f"jacobi_f_{m}"], INTERP[f"jacobi_out_{m}_mu{mu}"]
        got = orc.jacobirelaxation(v, fh, mu)
        assert np.abs(got - want).max() <= 4e-7 * mu * np.abs(want).max(), (m, mu, np.abs(got - want).max())
        # and it is NOT some other smoother: one sweep moves v by far more than the tolerance
        assert np.abs(orc.jacobirelaxation(v, fh, mu + 1) - want).max() > 1e-3
        # the oracle's reference-STRUCTURED path (CSR SpMV + scal/add passes, bench.py's CPU baseline A) reproduces the
        # reference's function bit for bit
        lvl = int(np.log2(m + 1))
        h = orc.csr_build(lvl, np.float32)
        try:
            assert np.array_equal(orc.csr_jacobirelaxation(h, v, fh, mu), want), (m, mu)
        finally:
            orc.csr_free(h, np.float32)


def test_jacobirelaxation_live_reference(orc, ref):
    m, mu = 23, 4
    rp, ci, va = _intended_a_lu(m)
    v = np.random.default_rng(11).uniform(-1, 1, m * m).astype(np.float32)
    fh = (1e-2 * np.random.default_rng(12).uniform(-1, 1, m * m)).astype(np.float32)
    out = v.copy()
    ref.ref_jacobirelaxation_with(m * m, P(rp), P(ci), P(va), P(out), P(fh), mu)
    assert np.abs(orc.jacobirelaxation(v, fh, mu) - out).max() <= 4e-7 * mu * np.abs(out).max()


def test_restriction_weight_as_written_is_zero_and_adjoint_pins_the_stencil(orc):
    """E2: `(1 / 16)` at P:539 is integer 0, so the reference's restriction2d returns zeros and
    cannot pin the stencil directly.  The stencil is pinned through P = 4 R^T against the
    interpolation that IS pinned above: <P x, y> == 4 <x, R_{1/16} y> for random x, y."""
    assert PINS["restriction2d_as_written_max_abs"] == 0.0
    m, n = 15, 31
    rng = np.random.default_rng(3)
    for _ in range(5):
        x, y = rng.uniform(-1, 1, m * m), rng.uniform(-1, 1, n * n)
        lhs = np.dot(orc.interpolation2d(x), y)
        rhs = 4.0 * np.dot(x, orc.restriction2d(y, w=1.0 / 16.0))
        assert lhs == pytest.approx(rhs, rel=1e-13)
    # and entry by entry: R = P^T / 4 as matrices
    Pm = np.stack([orc.interpolation2d(np.eye(1, m * m, k).ravel()) for k in range(m * m)], axis=1)
    Rm = np.stack([orc.restriction2d(np.eye(1, n * n, k).ravel(), w=1.0 / 16.0) for k in range(n * n)], axis=1)
    assert np.array_equal(Rm, Pm.T / 4.0)


def test_globalforcefunction_matches_reference_up_to_the_E3_sign(orc):
    """P:283-335 at the reference's finest level (10): every interior node gets |f h^2| = 4/2^20;
    as written the sign is negative (clockwise triangles, E3)."""
    L = PINS["finest_level"]
    b = orc.globalforcefunction(L, 4.0, np.float32)
    assert PINS["globalforce_size"] == b.size
    assert PINS["globalforce_min"] == PINS["globalforce_max"] == -float(b[0])


def test_assembled_operator_facts_E1_E3():
    """P:200-281 + coo_to_csr P:55-116: exact COO sums are +1 per off-diagonal coupling and -4
    per diagonal (A = -K, E3); the int32 accumulator of P:93 truncates them to 0 and -2 (E1)."""
    for lvl in (2, 3, 4):
        n = ((1 << lvl) - 1) ** 2
        side = (1 << lvl) - 1
        lu, d = PINS["csr"][f"L{lvl}_lu"], PINS["csr"][f"L{lvl}_d"]
        assert lu["min"] == lu["max"] == 0.0 and d["min"] == d["max"] == -2.0 and d["nnz"] == n
        assert d["coo_sum"] == -4.0 * n
        assert lu["coo_sum"] == 2 * 2 * side * (side - 1)      # one +1 per directed grid edge
        assert lu["max_row_nnz"] == 12                          # un-merged duplicates (E1b)


def test_cycle_call_structure_matches_oracle_cycle_shape():
    """vcyclemultigrid P:575-627 as executed by the reference: per non-coarsest level mu1+mu2
    sweeps, 1 residual (2 gemv + add + sub), 1 correction add; coarsest level mu1+mu2 sweeps.
    One sweep = 1 gemv + 2 scal + 2 add (P:138-142).  This is the shape oracle.vcyclemultigrid
    and Ctx::cycle_rec implement."""
    mu0, mu1, mu2 = PINS["mu0_mu1_mu2"]
    for lvl in (8, 9):
        nlev = lvl - PINS["coarsest_level"] + 1
        sweeps = nlev * (mu1 + mu2)
        c = PINS["vcycle_calls"][f"L{lvl}"]
        assert c["gemv"] == sweeps + 2 * (nlev - 1)
        assert c["scal"] == 2 * sweeps
        assert c["add"] == 2 * sweeps + 2 * (nlev - 1)
        assert c["sub"] == nlev - 1
    # fullmultigrid P:629-650: mu0+1 cycles per level on levels 7..10
    prog, cycles = PINS["program"], mu0 + 1
    visits = sum(cycles * (k + 1) for k in range(4))            # level 7+k is visited by cycles of levels >= itself
    visits = cycles * (4 + 3 + 2 + 1)
    residuals = cycles * (3 + 2 + 1)
    assert prog["scal"] == 2 * visits * (mu1 + mu2)
    assert prog["sub"] == residuals
    assert prog["gemv"] == visits * (mu1 + mu2) + 2 * residuals
    assert prog["add"] == 2 * visits * (mu1 + mu2) + 2 * residuals


def test_as_written_program_result_is_f_over_4():
    """with E1-E4 the program iterates v <- (1-w) v + (w/4) f_h on the finest
    grid only; after 620 sweeps in fp32 every entry equals f_h/4 = -2^-20 (it is not a Poisson
    solve, which is why parity is defined against the intended semantics)."""
    prog = PINS["program"]
    assert prog["size"] == 1023 * 1023 and prog["min"] == prog["max"]
    om = np.float32(2.0 / 3.0)
    c0, c1 = np.float32(1.0 - float(om)), np.float32(float(om) / 4.0)
    fh = np.float32(PINS["globalforce_min"])
    v = np.float32(0)
    for _ in range(620):
        v = np.float32(np.float32(c0 * v) + np.float32(c1 * fh))
    assert float(v) == prog["min"] == -2.0 ** -20


def test_live_reference_matches_fixture(ref):
    n = (1 << PINS["finest_level"]) - 1
    f = np.zeros(n * n, np.float32)
    assert ref.ref_globalforcefunction(P(f)) == PINS["globalforce_size"]
    assert float(f.min()) == PINS["globalforce_min"] and float(f.max()) == PINS["globalforce_max"]
    nn = 255
    u = np.zeros(nn * nn, np.float32)
    b = np.full(nn * nn, np.float32(-4.0 / 4 ** 8))
    ref.ref_counters_reset()
    ref.ref_vcyclemultigrid(8, P(u), P(b))
    c = (ctypes.c_longlong * 5)()
    ref.ref_counters(c)
    want = PINS["vcycle_calls"]["L8"]
    assert [c[0], c[1], c[2], c[3]] == [want["gemv"], want["scal"], want["add"], want["sub"]]
    assert float(u.min()) == want["u_min"] and float(u.max()) == want["u_max"]
