#!/bin/bash
# ONE 2-GPU gpurun call (charged 2x): parity + timing of the multi-rank paths written without a GPU.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round2_2gpu.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
echo "=== default multi-GPU parity" | tee -a $O/r02_steps2.log
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "2" > $O/r02_pytest_mgpu2.log 2>&1; tail -2 $O/r02_pytest_mgpu2.log
for knobs in "MGB200_GRAPH_DIST=1" "MGB200_COMM_AVOID=1" "MGB200_COMM_AVOID=1 MGB200_GRAPH_DIST=1" "MGB200_TILE=1 MGB200_ZERO_GUESS=1" \
             "MGB200_OVERLAP=1" "MGB200_OVERLAP=1 MGB200_GRAPH_DIST=1" "MGB200_CHAIN=1 MGB200_GRAPH_DIST=1"; do
    tag=$(echo $knobs | tr ' =' '__')
    echo "=== parity with $knobs" | tee -a $O/r02_steps2.log
    env $knobs timeout 300 $TR tests/mgpu_worker.py > $O/r02_mgpu_$tag.log 2>&1; tail -1 $O/r02_mgpu_$tag.log
done
for knobs in "X=1" "MGB200_GRAPH_DIST=1" "MGB200_COMM_AVOID=1" "MGB200_COMM_AVOID=1 MGB200_GRAPH_DIST=1" "MGB200_OVERLAP=1" \
             "MGB200_OVERLAP=1 MGB200_GRAPH_DIST=1" "MGB200_CHAIN=1 MGB200_GRAPH_DIST=1" "MGB200_CHAIN=1 MGB200_OVERLAP=1 MGB200_GRAPH_DIST=1"; do
    tag=$(echo $knobs | tr ' =' '__')
    echo "=== bench 16385^2 on 2 GPUs with $knobs" | tee -a $O/r02_steps2.log
    env $knobs timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/r02_bench_n2_$tag.json 2> $O/r02_bench_n2_$tag.err
    python -c "
import json
d = json.loads([l for l in open('$O/r02_bench_n2_$tag.json') if l.startswith('{')][-1])
print('$tag', 'ms', round(d['ms_per_step'], 4), 'n1 same workload ms', d.get('n1_same_workload', {}).get('ms_per_step'))" || echo "$tag FAILED"
done
