"""sampled load vectors / Dirichlet data (generalised globalforcefunction, P:283-335) and
the v2 call shape `ProblemVar` + `multigrid_solver` (M:16-26, M:175-197).  Host-side checks run everywhere; the
solver checks are `gpu` tests (they also run on the CPU suite through the emulated library,
tests/test_emulated_library.py)."""
import numpy as np
import pytest

import oracle
from conftest import assert_bitwise, rand_vec


def test_load_vector_constant_f_is_the_reference_rhs(mgb, orc):
    for level in (1, 3, 6, 8):
        assert_bitwise(mgb.load_vector(level, 4.0), orc.globalforcefunction(level), f"b = 4 h^2 at level {level}")
        assert_bitwise(mgb.load_vector(level, lambda x, y: 4.0 + 0 * x), orc.globalforcefunction(level), "callable f")
    b32 = mgb.load_vector(5, 4.0, dtype=np.float32)
    assert b32.dtype == np.float32 and b32[0] == np.float32(4.0 / 1024)


def test_load_vector_layout_and_dirichlet_folding(mgb):
    level, N = 3, 8
    n = N - 1
    b = mgb.load_vector(level, lambda x, y: x + 10 * y).reshape(n, n)            # row <-> y (P:227-228)
    assert b[2, 4] == pytest.approx((5 / N + 10 * 3 / N) / N ** 2, rel=1e-15)
    g = lambda x, y: 1 + x + 100 * y                                             # noqa: E731
    d = (mgb.load_vector(level, 0.0, g)).reshape(n, n)
    assert d[3, 3] == 0                                                          # interior rows see no boundary
    assert d[0, 2] == pytest.approx(g(3 / N, 0.0))                               # bottom ring
    assert d[n - 1, 2] == pytest.approx(g(3 / N, 1.0))                           # top ring
    assert d[2, 0] == pytest.approx(g(0.0, 3 / N)) and d[2, n - 1] == pytest.approx(g(1.0, 3 / N))
    assert d[0, 0] == pytest.approx(g(1 / N, 0.0) + g(0.0, 1 / N))               # corner node: two ring neighbours


def _oracle_v2(orc, obj, p):
    """M:175-191 restated with the oracle's operators."""
    b = dict(obj.b_dict)
    for l in range(obj.finest_level - 1, obj.coarsest_level - 1, -1):
        if l not in b:
            b[l] = orc.restriction2d(b[l + 1])
    u = np.zeros_like(b[obj.coarsest_level])
    for _ in range(obj.mu0 + 1):
        u = orc.vcyclemultigrid(u, b[obj.coarsest_level], p)
    for l in range(obj.coarsest_level + 1, obj.finest_level + 1):
        u = orc.interpolation2d(u)
        for _ in range(obj.mu0 + 1):
            u = orc.vcyclemultigrid(u, b[l], p)
    return u


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("smoother", ["jacobi", "rbgs"])
@pytest.mark.parametrize("coarse", ["exact", "sweeps"])
def test_multigrid_solver_v2_shape_bitwise(mgb, orc, dtype, smoother, coarse):
    """multigrid_solver(ProblemVar): per-level load vectors where given (M:183), restricted ones elsewhere (P:641)."""
    top = 7
    obj = mgb.ProblemVar(finest_level=top, coarsest_level=2, mu0=2, mu1=1, mu2=1, smoother=smoother, dtype=dtype, coarse_solver=coarse)
    obj.b_dict[top] = rand_vec(top, dtype, 61, 1e-3)
    obj.b_dict[top - 2] = rand_vec(top - 2, dtype, 62, 1e-3)         # an independently assembled coarse load vector
    p = oracle.Params(coarsest_level=2, nu1=1, nu2=1, smoother=1 if smoother == "rbgs" else 0, nthreads=2,
                      coarse_exact=1 if coarse == "exact" else 0)
    assert_bitwise(mgb.multigrid_solver(obj), _oracle_v2(orc, obj, p), "multigrid_solver")


@pytest.mark.gpu
def test_dirichlet_problem_reproduces_a_harmonic_quadratic(mgb):
    """-Lap u = 0, u = x^2 - y^2 on the boundary: the 5-point stencil is exact for quadratics, so the discrete
    solution equals the exact one at the nodes (to the solver tolerance).  Also -Lap u = -4 with u = x^2 + y^2."""
    level = 7
    x, y = mgb.problem.node_coordinates(level)
    for f, g in ((0.0, lambda x, y: x * x - y * y), (-4.0, lambda x, y: x * x + y * y)):
        with mgb.Multigrid(level) as mg:
            mg.set_rhs(level, mgb.load_vector(level, f, g))
            mg.zero_u(level)
            k, rel, _ = mg.solve(1e-13, 60)
            assert rel <= 1e-13 and k < 30
            u = mg.get_u(level).reshape(x.shape)
        assert np.abs(u - g(x, y)).max() < 1e-11
