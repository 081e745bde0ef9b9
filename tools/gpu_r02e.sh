#!/bin/bash
# N-GPU call (N = $1): the new N>1 bench arm (configs[2]: 16385^2 RB-GS; strong_scaling + mgpu_parity) with the schedule knobs
set -u
N=${1:-2}
mkdir -p gpurun_out; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
show() { python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/{tag}.json") if l.startswith("{")][-1])
    ss = d.get("strong_scaling") or {}
    print(f"{tag:44s} ms {d['ms_per_step']:.4f} iso {d['isolated_cycle_ms']:.4f} launches {d['gpu_launches']} | n1 {ss.get('n1_ms_per_step')} eff {ss.get('efficiency')} iso_eff {ss.get('isolated_efficiency')} parity {ss.get('mgpu_parity')} | pre {d['roofline']['ms_per_launch']:.4f} frac {d['roofline']['frac']:.3f} | solve {d['solve'].get('cycles')} {d['solve'].get('ms')}")
    if d.get("extra"): print("     extra", d["extra"])
except Exception as ex:
    print(tag, "FAILED", ex, open(f"gpurun_out/{tag}.err").read()[-1500:])
PY
}
run() { tag=r02e_n${N}_$1; shift; env "$@" timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e $EXTRA > $O/$tag.json 2> $O/$tag.err; show $tag; }
EXTRA="" run default X=1
EXTRA="--no-n1 --no-extra" run lazy_eager MGB200_COMM_AVOID=0 MGB200_GRAPH_DIST=0
EXTRA="--no-n1 --no-extra" run lazy_graph MGB200_COMM_AVOID=0
EXTRA="--no-n1 --no-extra" run ca_eager MGB200_GRAPH_DIST=0
