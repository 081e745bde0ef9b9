#!/bin/bash
# 1-GPU call after the round-2 cleanup: default suite (now incl. zero-guess / visit chains / full-size configs), benches.
set -u
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $O/r02b_pytest.log 2>&1; echo "rc=$?" >> $O/r02b_pytest.log; tail -15 $O/r02b_pytest.log
b() { tag=$1; shift; timeout 300 python bench.py --no-cpu --no-e2e "$@" > $O/r02b_bench_$tag.json 2> $O/r02b_bench_$tag.err; python - $tag <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02b_bench_{tag}.json") if l.startswith("{")][-1])
    k = d["roofline"]["kernels"]
    print(f"{tag:24s} cycle {d['ms_per_step']*1e3:8.1f} us iso {d.get('isolated_cycle_ms', 0)*1e3:8.1f} launches {d['gpu_launches']:4d} " +
          " ".join(f"{n[:8]}:{v['ms']*1e3:.0f}us/{v['frac_of_peak']:.2f}" for n, v in k.items()))
    print("   levels", {a: round(b * 1e3) for a, b in d["roofline"]["cycle_ms_from_level_down"].items()}, "solve", d["solve"]["cycles"], round(d["solve"]["ms"], 2))
except Exception as ex:
    print(tag, "FAILED", ex, open(f"gpurun_out/r02b_bench_{tag}.err").read()[-800:])
PY
}
b default
MGB200_CHAIN=0 b nochain
b rbgs_L12 --smoother rbgs
b rbgs_L14 --smoother rbgs --level 14
b W_L13 --level 13 --gamma 2
MGB200_CHAIN=0 b W_L13_nochain --level 13 --gamma 2
b jac_L14 --level 14
