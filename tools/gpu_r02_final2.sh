#!/bin/bash
# FINAL N-GPU call (N = $1): exactly the driver's command for the scaling run, ONE run, under its own short timeout
set -u
N=${1:-4}
mkdir -p gpurun_out; O=gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_final_bench_n$N.json 2> $O/r02_final_bench_n$N.err
echo "rc=$?"
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02_final_bench_n{N}.json") if l.startswith("{")][-1])
    ss = d.get("strong_scaling") or {}
    print(f"N={N} ms {d['ms_per_step']:.4f} iso {d['isolated_cycle_ms']:.4f} | n1 {ss.get('n1_ms_per_step')} eff {ss.get('efficiency')} parity {ss.get('mgpu_parity')} | pre {d['roofline']['ms_per_launch']:.4f} frac {d['roofline']['frac']:.3f} | e2e {d['e2e']['ms_per_step']:.2f} ms | solve {d['solve'].get('cycles')} {d['solve'].get('ms')}")
    print("   extra", d.get("extra"))
except Exception as ex:
    print("FAILED", ex, open(f"gpurun_out/r02_final_bench_n{N}.err").read()[-1500:])
PY
