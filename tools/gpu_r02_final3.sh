#!/bin/bash
# last check of HEAD on one GPU: the whole GPU suite and the default bench line (what the driver runs at round end)
set -u
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02_head_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02_head_pytest_gpu.log; tail -4 $O/r02_head_pytest_gpu.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_head_bench.json 2> $O/r02_head_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_head_bench.json") if l.startswith("{")][-1])
r = d["roofline"]
print("value", d["value"], "ms", d["ms_per_step"], "iso", d["isolated_cycle_ms"], "launches", d["gpu_launches"])
print("roofline", r["kernel"][:60], "frac", r["frac"], "share", r["share_of_step"], "traffic", r["traffic"], "two-launch", (r["two_launch_equivalent"] or {}).get("frac"))
print("e2e", d["e2e"]["ms_per_step"], "pageable", d["e2e"]["pageable"]["ms_per_step"], "solve", d["solve"]["ms"], d["solve"]["overhead_vs_isolated_cycle"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "clocks", d["clocks"])
PY
timeout 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 | cut -c1-400
