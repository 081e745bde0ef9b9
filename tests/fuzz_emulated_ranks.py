"""Randomised multi-rank parity check on the CPU emulation build: random levels, agglomeration levels, smoothers, sweep
counts, cycle index, schedule knobs and call sequences on row slabs, every rank's rows against the single-domain oracle.
    MGB200_EMU_DIR=/tmp/x python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 \
        --master-port <free> tests/fuzz_emulated_ranks.py [seconds] [seed]
TEST TOOLING (every rank draws the same random numbers)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["MGB200_TEST_EMU"] = "1"
import numpy as np  # noqa: E402
import torch.distributed as dist  # noqa: E402

import conftest  # noqa: E402

conftest.use_emulated_library()
import mgb200  # noqa: E402
import oracle  # noqa: E402

mgdist = __import__("importlib").import_module("multigrid_nikhil_c-_b200.dist")
KNOBS = ("MGB200_COMM_AVOID", "MGB200_GRAPH_DIST", "MGB200_CHAIN", "MGB200_ZERO_GUESS")


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dist.init_process_group("gloo")
    os.environ["LOCAL_RANK"] = "0"
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(seed)
    o = oracle.get()
    t0, n = time.time(), 0
    while True:
        stop = [time.time() - t0 > budget]
        dist.broadcast_object_list(stop, src=0)
        if stop[0]:
            break
        level = int(rng.integers(6, 10))
        min_aggl = 2
        while (1 << (min_aggl + 1)) // world < 8:
            min_aggl += 1
        aggl = int(rng.integers(min_aggl, level))
        dtype = [np.float64, np.float32][int(rng.integers(0, 2))]
        smoother = ["jacobi", "rbgs"][int(rng.integers(0, 2))]
        nu1, nu2, gamma = int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.integers(1, 3))
        flags = dict(graph=bool(rng.integers(0, 2)), fused=bool(rng.integers(0, 4) > 0), coarse_tail=bool(rng.integers(0, 2)))
        env = {k: str(int(rng.integers(0, 2))) for k in KNOBS}
        if os.environ.get("FUZZ_DEFAULT_ONLY") == "1":      # the default configuration only (what the GPU suite runs)
            env = {k: "1" for k in KNOBS}   # the library defaults
        os.environ.update(env)
        cfg = dict(level=level, aggl=aggl, dtype=np.dtype(dtype).name, smoother=smoother, nu1=nu1, nu2=nu2, gamma=gamma, **flags, **env)
        m = (1 << level) - 1
        x = rng.uniform(-1, 1, m * m).astype(dtype)
        b = (1e-3 * rng.uniform(-1, 1, m * m)).astype(dtype)
        sid = 1 if smoother == "rbgs" else 0
        p = oracle.Params(nu1=nu1, nu2=nu2, gamma=gamma, smoother=sid, nthreads=1)
        ops = [int(v) for v in rng.integers(0, 7, size=int(rng.integers(2, 6)))]
        cnts = [int(v) for v in rng.integers(2, 4, size=len(ops))]
        try:
            mg = mgdist.create(level, dtype=dtype, smoother=smoother, agglomerate_level=aggl, **flags)
            mg.set_u(level, x)
            mg.set_rhs(level, b)
            want = x
            for op, cnt in zip(ops, cnts):
                if op == 0:
                    mg.cycle(level, nu1, nu2, gamma)
                    want = o.vcyclemultigrid(want, b, p)
                elif op == 1:
                    mg.cycles(cnt, level, nu1, nu2, gamma)
                    for _ in range(cnt):
                        want = o.vcyclemultigrid(want, b, p)
                elif op == 2:
                    mg.smooth(level, cnt)
                    want = o.jacobirelaxation(want, b, cnt) if sid == 0 else o.rbgs(want, b, cnt)
                elif op == 3:
                    nrm = mg.residual(level, norm=True)
                    r = o.residual(want, b)
                    if not np.array_equal(mgdist.owned_slice(mg, level, mg.get_r(level)), mgdist.owned_slice(mg, level, r)):
                        raise AssertionError("residual differs")
                    if abs(nrm - o.norm2(r)) > 1e-5 * max(o.norm2(r), 1e-30):
                        raise AssertionError("norm differs")
                elif op == 6:                    # tolerance loop: the norm comes out of the cycle's last kernel where that applies
                    ps = oracle.Params(nu1=max(nu1, 1), nu2=max(nu2, 1), gamma=gamma, smoother=sid, nthreads=1)
                    k, rel, hist = mg.solve(1e-6, cnt, ps.nu1, ps.nu2, gamma)
                    want, ko, ho = o.solve(want, b, 1e-6, cnt, ps)
                    if k != ko or not np.allclose(hist, ho, rtol=1e-9, atol=0):
                        raise AssertionError(f"solve differs: {k} vs {ko} cycles, {hist} vs {ho}")
                elif op == 4:
                    mg.set_u(level, want)        # host round trip: halos become valid again
                else:
                    pf = oracle.Params(nu1=max(nu1, 1), nu2=max(nu2, 1), smoother=sid, nthreads=1)
                    got = mg.fullmultigrid(b, 1, pf.nu1, pf.nu2)
                    if not np.array_equal(mgdist.owned_slice(mg, level, got), mgdist.owned_slice(mg, level, o.fullmultigrid(b, 1, pf))):
                        raise AssertionError("fullmultigrid differs")
                    mg.set_u(level, want)
                    mg.set_rhs(level, b)
                if not np.array_equal(mgdist.owned_slice(mg, level, mg.get_u(level)), mgdist.owned_slice(mg, level, want)):
                    raise AssertionError(f"iterate differs after op {op} (count {cnt})")
            mg.close()
        except Exception as ex:  # noqa: BLE001
            print(f"rank {rank} FAIL", cfg, "ops", ops, cnts, "->", repr(ex), flush=True)
            os._exit(1)
        n += 1
    if rank == 0:
        print(f"rank fuzz OK: {n} random configurations on {world} ranks in {time.time() - t0:.0f} s (seed {seed})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
