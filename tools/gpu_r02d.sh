#!/bin/bash
# 1-GPU call: tail variants (single-warp small levels: off / <=3 / <=4), solve with the norm folded into POST,
# ncu source-level capture of the red-black PRE kernel
set -u
mkdir -p gpurun_out; O=gpurun_out
L=multigrid_nikhil_c-_b200/lib
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "solve or combinations or fullmultigrid or synthetic" > $O/r02d_pytest.log 2>&1; echo "rc=$?" >> $O/r02d_pytest.log; tail -4 $O/r02d_pytest.log
for v in _tw0 _tw3 ""; do
  MGB200_LIB=$PWD/$L/libmgb200$v.so python - "$v" <<'PY'
import sys, statistics, numpy as np
sys.path.insert(0, '.')
import mgb200
tag = sys.argv[1] or "_tw4"
out = []
for level, gamma in ((6, 1), (6, 2), (9, 1), (9, 2), (12, 1), (13, 2)):
    with mgb200.Multigrid(level) as mg:
        mg.force_synthetic(1234); mg.zero_u(level)
        mg.time_cycle(level, 2, 2, gamma, 3)
        t = statistics.median([mg.time_cycle(level, 2, 2, gamma, 1) for _ in range(15)])
        out.append(f"L{level}g{gamma}:{t*1e3:.1f}us")
print(f"tail{tag:6s}", " ".join(out))
PY
done
python - <<'PY'
import sys, time, statistics
sys.path.insert(0, '.')
import mgb200
for sm in ("jacobi", "rbgs"):
    with mgb200.Multigrid(12, smoother=sm) as mg:
        mg.force_synthetic(1234); mg.zero_u(12)
        mg.solve(1e-8, 3); mg.zero_u(12); mg.sync()
        t0 = time.perf_counter(); k, rel, h = mg.solve(1e-8, 40); mg.sync(); ms = (time.perf_counter() - t0) * 1e3
        mg.time_cycle(12, 2, 2, 1, 1)
        iso = statistics.median([mg.time_cycle(12, 2, 2, 1, 1) for _ in range(15)])
        print(f"solve {sm}: {k} cycles {ms:.3f} ms = {ms/k*1e3:.1f} us/cycle, isolated cycle {iso*1e3:.1f} us, overhead {ms/k/iso-1:+.1%}, relres {rel:.3e}")
PY
CMD="python bench.py --no-cpu --no-e2e --smoother rbgs --steps 2 --warmup 3"
MGB200_AUTOTUNE=0 $CMD > $O/r02d_plain.log 2>&1 && MGB200_AUTOTUNE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_stream<double, 4, 1" -s 3 -c 1 -o $O/r02d_prof_rbgs_pre $CMD > $O/r02d_ncu.log 2>&1
tail -2 $O/r02d_ncu.log
