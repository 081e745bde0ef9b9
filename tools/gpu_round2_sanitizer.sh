#!/bin/bash
# A separate 1-GPU gpurun call: compute-sanitizer memcheck (ONE tool per call, B200_PROFILING.md) on one small cycle test
# of the default path and of the opt-in kernels.  Run only after tools/gpu_round2_1gpu.sh passed (the same commands
# must have exited 0 without the tool first).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round2_sanitizer.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
SEL='cycles_bitwise_all_flag_combinations and 7-jacobi-2-2-1-1-float64'
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL" > $O/r02_san_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/r02_san_plain.log; exit 1; }
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL" > $O/r02_san_memcheck_default.log 2>&1
echo "memcheck default rc=$?"; tail -3 $O/r02_san_memcheck_default.log
for knobs in "MGB200_ZERO_GUESS=1 MGB200_CHAIN=1" "MGB200_TILE=1" "MGB200_CTAIL=1"; do
    tag=$(echo $knobs | tr ' =' '__')
    env $knobs timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL" \
        > $O/r02_san_memcheck_$tag.log 2>&1
    echo "memcheck $knobs rc=$?"; tail -2 $O/r02_san_memcheck_$tag.log
done
