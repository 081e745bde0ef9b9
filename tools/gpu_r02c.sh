#!/bin/bash
# 1-GPU call: ring-depth experiment (3 builds of the same sources) + ncu source-level capture of the coarse tail
set -u
mkdir -p gpurun_out; O=gpurun_out
L=multigrid_nikhil_c-_b200/lib
b() { tag=$1; shift; timeout 300 python bench.py --no-cpu --no-e2e "$@" > $O/r02c_bench_$tag.json 2> $O/r02c_bench_$tag.err; python - $tag <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02c_bench_{tag}.json") if l.startswith("{")][-1])
    k = d["roofline"]["kernels"]
    print(f"{tag:26s} cycle {d['ms_per_step']*1e3:8.1f} us iso {d.get('isolated_cycle_ms', 0)*1e3:8.1f} " +
          " ".join(f"{n[:8]}:{v['ms']*1e3:.0f}us/{v['frac_of_peak']:.2f}" for n, v in k.items()))
except Exception as ex:
    print(tag, "FAILED", ex, open(f"gpurun_out/r02c_bench_{tag}.err").read()[-800:])
PY
}
for v in "" _deep3 _deep1; do
  export MGB200_LIB=$PWD/$L/libmgb200$v.so
  b jac_L12$v
  b rbgs_L12$v --smoother rbgs
  b rbgs_L14$v --smoother rbgs --level 14
  MGB200_STREAM_OCC=6 b rbgs_L14_occ6$v --smoother rbgs --level 14
done
export MGB200_LIB=$PWD/$L/libmgb200_deep1.so
MGB200_STREAM_OCC=6 b jac_L12_occ6_deep1
unset MGB200_LIB
CMD="python bench.py --no-cpu --no-e2e --level 6 --steps 3 --warmup 3"
$CMD > $O/r02c_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tail -s 6 -c 1 -o $O/r02c_prof_tail $CMD > $O/r02c_ncu_tail.log 2>&1
tail -3 $O/r02c_ncu_tail.log
