#!/bin/bash
# 1-GPU call: device-loop solve (conditional WHILE graph) -- parity tests, then timing against the host loop
set -u
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "device_side_loop or solve_257" > $O/r02g_pytest.log 2>&1; echo "rc=$?" >> $O/r02g_pytest.log; tail -15 $O/r02g_pytest.log
for sg in 0 1; do
MGB200_SOLVE_GRAPH=$sg python - $sg <<'PY'
import sys, time, statistics
sys.path.insert(0, '.')
import mgb200
for level, sm in ((12, "jacobi"), (12, "rbgs"), (10, "jacobi")):
    with mgb200.Multigrid(level, smoother=sm) as mg:
        mg.force_synthetic(1234); mg.zero_u(level)
        mg.solve(1e-8, 40); 
        ts = []
        for _ in range(5):
            mg.zero_u(level); mg.sync()
            t0 = time.perf_counter(); k, rel, h = mg.solve(1e-8, 40); mg.sync(); ts.append((time.perf_counter() - t0) * 1e3)
        ms = statistics.median(ts)
        mg.time_cycle(level, 2, 2, 1, 1)
        iso = statistics.median([mg.time_cycle(level, 2, 2, 1, 1) for _ in range(15)])
        print(f"SOLVE_GRAPH={sys.argv[1]} L{level} {sm}: {k} cycles {ms:.3f} ms = {ms/k*1e3:.1f} us/cycle, isolated cycle {iso*1e3:.1f} us, overhead {ms/k/iso-1:+.1%}, relres {rel:.3e}", flush=True)
PY
done
