#!/bin/bash
# ONE 1-GPU gpurun call that verifies and times everything written without a GPU (see NOTES_NEXT_ROUND.md):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round2_1gpu.sh'
# Every step has its own timeout; outputs land in gpurun_out/r02_*.  Nothing here runs under ncu unless the same
# command has just exited 0 without it.
set -u
mkdir -p gpurun_out
O=gpurun_out
step() { echo "=== $*" | tee -a $O/r02_steps.log; }

step "default GPU suite"
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02_pytest_default.log 2>&1; echo "rc=$?" >> $O/r02_pytest_default.log
tail -3 $O/r02_pytest_default.log

step "opt-in paths (tile kernels, zero-guess chain, cluster tail)"
MGB200_TEST_OPTIN=1 timeout 600 python -m pytest tests/test_optin_gpu.py -m gpu -q -k "tile_kernels_cycles or zero_guess or cluster_tail or visit_chain" \
    > $O/r02_pytest_optin.log 2>&1; echo "rc=$?" >> $O/r02_pytest_optin.log
tail -3 $O/r02_pytest_optin.log

step "TMA variant of the streaming kernels, alone and under a short timeout (a wrong mbarrier transaction count would hang)"
MGB200_TEST_OPTIN=1 timeout 180 python -m pytest tests/test_optin_gpu.py -m gpu -x -q -k "tma_streaming" > $O/r02_pytest_tma.log 2>&1; echo "rc=$?" >> $O/r02_pytest_tma.log
tail -3 $O/r02_pytest_tma.log

run_bench() {   # tag, env assignments...
    local tag=$1; shift
    env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 > $O/r02_bench_$tag.json 2> $O/r02_bench_$tag.err
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02_bench_{tag}.json") if l.startswith("{")][-1])
    lv = d["roofline"]["cycle_ms_from_level_down"]
    print(f"{tag:28s} cycle {d['ms_per_step']*1e3:7.1f} us  launches {d['gpu_launches']:3d}  from-level-down(us): " +
          " ".join(f"L{k}:{v*1e3:.0f}" for k, v in lv.items()))
except Exception as ex:
    print(tag, "FAILED", ex)
PY
}
step "bench, one knob at a time (X=1 is a no-op placeholder)"
run_bench default X=1
run_bench zero_guess MGB200_ZERO_GUESS=1
run_bench tile MGB200_TILE=1
run_bench tile512 MGB200_TILE=1 MGB200_TILE_MAXN=512
run_bench ctail16 MGB200_CTAIL=1 MGB200_CTAIL_CTAS=16
run_bench ctail8 MGB200_CTAIL=1 MGB200_CTAIL_CTAS=8
run_bench zg_tile MGB200_ZERO_GUESS=1 MGB200_TILE=1
run_bench zg_ctail16 MGB200_ZERO_GUESS=1 MGB200_CTAIL=1
run_bench zg_tile_ctail16 MGB200_ZERO_GUESS=1 MGB200_TILE=1 MGB200_CTAIL=1
grep -q "rc=0" $O/r02_pytest_tma.log && run_bench tma MGB200_TMA=1
run_bench chain MGB200_CHAIN=1
run_bench chain_zg MGB200_CHAIN=1 MGB200_ZERO_GUESS=1
run_bench chain_zg_tile_ctail16 MGB200_CHAIN=1 MGB200_ZERO_GUESS=1 MGB200_TILE=1 MGB200_CTAIL=1
step "other BASELINE configs on one GPU"
timeout 300 python bench.py --no-cpu --no-e2e --level 13 --gamma 2 > $O/r02_bench_cfg4_W_8193.json 2>> $O/r02_steps.log
timeout 300 python bench.py --no-cpu --no-e2e --level 14 --smoother rbgs > $O/r02_bench_cfg3_rbgs_16385_n1.json 2>> $O/r02_steps.log
timeout 300 python bench.py --micro --dtype f32 --level 15 > $O/r02_bench_cfg5_micro_f32.json 2>> $O/r02_steps.log

step "full default bench line (with CPU baseline and e2e)"
timeout 600 python bench.py > $O/r02_bench_full.json 2> $O/r02_bench_full.err

step "ncu: launch list of two V-cycles, then full capture of the streaming kernels"
CMD="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3"
$CMD > $O/r02_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv $CMD > $O/r02_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 40 -c 4 -o $O/r02_prof_stream $CMD > $O/r02_ncu2.log 2>&1
step done
