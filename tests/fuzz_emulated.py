"""Randomised parity check of the whole library against the oracle on the CPU emulation build (tests/host_emul):
random levels, coarsest levels, sweep counts, cycle index, smoother, dtype, flags and call sequences.
    python tests/fuzz_emulated.py [seconds] [seed]
TEST TOOLING; prints the failing configuration and exits 1 on the first mismatch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["MGB200_TEST_EMU"] = "1"
import numpy as np  # noqa: E402

import conftest  # noqa: E402

conftest.use_emulated_library()
import mgb200  # noqa: E402
import oracle  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    o = oracle.get()
    t0, n = time.time(), 0
    while time.time() - t0 < budget:
        level = int(rng.integers(1, 9))
        coarsest = int(rng.integers(1, level + 1))
        dtype = [np.float64, np.float32][int(rng.integers(0, 2))]
        smoother = ["jacobi", "rbgs"][int(rng.integers(0, 2))]
        nu1, nu2 = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        gamma = int(rng.integers(1, 4))
        flags = dict(graph=bool(rng.integers(0, 2)), fused=bool(rng.integers(0, 2)), coarse_tail=bool(rng.integers(0, 2)))
        exact = bool(rng.integers(0, 4) == 0)          # MG_COARSE_EXACT: direct solve on the coarsest level (M:63-72)
        flags["coarse_solver"] = "exact" if exact else "sweeps"
        env = {k: str(int(rng.integers(0, 2))) for k in ("MGB200_ZERO_GUESS", "MGB200_CHAIN")}
        if os.environ.get("FUZZ_DEFAULT_ONLY") == "1":
            env = {k: "1" for k in env}   # the library defaults
        os.environ.update(env)
        cfg = dict(level=level, coarsest=coarsest, dtype=np.dtype(dtype).name, smoother=smoother, nu1=nu1, nu2=nu2, gamma=gamma,
                   **flags, **env)
        m = (1 << level) - 1
        x = rng.uniform(-1, 1, m * m).astype(dtype)
        b = (1e-3 * rng.uniform(-1, 1, m * m)).astype(dtype)
        p = oracle.Params(coarsest_level=coarsest, nu1=nu1, nu2=nu2, gamma=gamma, smoother=1 if smoother == "rbgs" else 0, nthreads=1,
                          coarse_exact=int(exact))
        try:
            with mgb200.Multigrid(level, coarsest_level=coarsest, dtype=dtype, smoother=smoother, **flags) as mg:
                mg.set_u(level, x)
                mg.set_rhs(level, b)
                want = x
                for k in range(int(rng.integers(1, 4))):
                    cnt = int(rng.integers(1, 4))                  # 1: mg_cycle, > 1: mg_cycles (visit chains)
                    if cnt == 1:
                        mg.cycle(level, nu1, nu2, gamma)
                    else:
                        mg.cycles(cnt, level, nu1, nu2, gamma)
                    for _ in range(cnt):
                        want = o.vcyclemultigrid(want, b, p)
                    if not np.array_equal(mg.get_u(level), want):
                        raise AssertionError(f"cycle {k + 1} (x{cnt}) differs, max {np.abs(mg.get_u(level) - want).max():.3e}")
                # an operator sequence after the cycles, then another cycle (state bookkeeping across API calls)
                nrm = mg.residual(level, norm=True)
                r = o.residual(want, b)
                if not np.array_equal(mg.get_r(level), r):
                    raise AssertionError("residual after cycles differs")
                if abs(nrm - o.norm2(r)) > 1e-6 * max(o.norm2(r), 1e-30):
                    raise AssertionError("norm differs")
                if level > coarsest:
                    mg.restrict(level)
                    if not np.array_equal(mg.get_rhs(level - 1), o.restriction2d(r)) or mg.get_u(level - 1).any():
                        raise AssertionError("restrict after cycles differs")
                mg.smooth(level, 1 + nu1)
                want = o.jacobirelaxation(want, b, 1 + nu1) if smoother == "jacobi" else o.rbgs(want, b, 1 + nu1)
                mg.cycle(level, max(nu1, 1), nu2, gamma)
                p2 = oracle.Params(coarsest_level=coarsest, nu1=max(nu1, 1), nu2=nu2, gamma=gamma, smoother=p.smoother, nthreads=1,
                                   coarse_exact=int(exact))
                want = o.vcyclemultigrid(want, b, p2)
                if not np.array_equal(mg.get_u(level), want):
                    raise AssertionError("cycle after operator calls differs")
                if rng.integers(0, 2):      # tolerance loop (norm folded into the cycle's last kernel where it applies)
                    ps = oracle.Params(coarsest_level=coarsest, nu1=max(nu1, 1), nu2=max(nu2, 1), gamma=gamma, smoother=p.smoother,
                                       nthreads=1, coarse_exact=int(exact))
                    mc = int(rng.integers(1, 5))
                    k, rel, hist = mg.solve(1e-6, mc, ps.nu1, ps.nu2, gamma)
                    us, ko, ho = o.solve(want, b, 1e-6, mc, ps)
                    if k != ko or not np.allclose(hist, ho, rtol=1e-9, atol=0) or not np.array_equal(mg.get_u(level), us):
                        raise AssertionError(f"solve differs: {k} vs {ko} cycles")
                    want = us
                if rng.integers(0, 2):
                    pf = oracle.Params(coarsest_level=coarsest, nu1=max(nu1, 1), nu2=max(nu2, 1), smoother=p.smoother, nthreads=1,
                                       coarse_exact=int(exact))
                    cyc = 1 + int(rng.integers(0, 2))
                    if not np.array_equal(mg.fullmultigrid(b, cyc, pf.nu1, pf.nu2), o.fullmultigrid(b, cyc, pf)):
                        raise AssertionError("fullmultigrid differs")
        except Exception as ex:  # noqa: BLE001
            print("FAIL", cfg, "->", repr(ex), flush=True)
            sys.exit(1)
        n += 1
    print(f"fuzz OK: {n} random configurations in {time.time() - t0:.0f} s (seed {seed})")


if __name__ == "__main__":
    main()
