#!/bin/bash
# 1-GPU call: ncu evidence for the shipped kernels, per kernel and per level (each command first runs clean without ncu)
set -u
mkdir -p gpurun_out; O=gpurun_out
# A/B: skewed pipeline (libmgb200.so) vs the previous kernel (libmgb200_noskew.so), each with and without programmatic
# dependent launches; per-kernel times on the finest level and cycle times (isolated and 10 chained)
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "not 4097 and not cpp_" > $O/r02f_pytest_skew.log 2>&1; echo "rc=$?" >> $O/r02f_pytest_skew.log; tail -3 $O/r02f_pytest_skew.log
for lib in libmgb200_noskew.so libmgb200.so; do for pdl in 0 1; do
MGB200_LIB=$PWD/multigrid_nikhil_c-_b200/lib/$lib MGB200_PDL=$pdl python - $lib $pdl <<'PY'
import os, sys, statistics
sys.path.insert(0, '.')
import mgb200
from mgb200 import capi
out = []
for level, gamma, sm in ((9, 1, "jacobi"), (12, 1, "jacobi"), (12, 1, "rbgs"), (14, 1, "rbgs"), (13, 2, "jacobi")):
    with mgb200.Multigrid(level, smoother=sm) as mg:
        mg.force_synthetic(1234); mg.zero_u(level)
        mg.time_cycle(level, 2, 2, gamma, 3)
        iso = statistics.median([mg.time_cycle(level, 2, 2, gamma, 1) for _ in range(11)])
        mg.time_cycle(level, 2, 2, gamma, 10)
        ch = statistics.median([mg.time_cycle(level, 2, 2, gamma, 10) for _ in range(5)]) / 10
        ops = []
        for name, op in (("pre", capi.MG_OP_PRE_FUSED), ("post", capi.MG_OP_POST_FUSED), ("chain", capi.MG_OP_POSTPRE_FUSED), ("sw2", capi.MG_OP_SMOOTH2)):
            mg.time_op(op, level, 2)
            ops.append(f"{name} {mg.time_op(op, level, 10) / 10 * 1e3:.0f}")
        out.append(f"L{level}g{gamma}{sm[0]}: iso {iso*1e3:.1f} chained {ch*1e3:.1f} [{' '.join(ops)}]")
print(f"{sys.argv[1]:22s} PDL={sys.argv[2]}", " | ".join(out), flush=True)
PY
done; done
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
cap() { tag=$1; shift; python tools/profile_ops.py "$@" > $O/r02f_$tag.plain.log 2>&1 && timeout 600 $NCU -k regex:"k_stream" -c 12 -o $O/r02f_$tag python tools/profile_ops.py "$@" > $O/r02f_$tag.ncu.log 2>&1; tail -1 $O/r02f_$tag.ncu.log; }
cap jac_L12 12 jacobi f64 pre post chain
cap jac_L11 11 jacobi f64 pre post chain
cap jac_L10 10 jacobi f64 pre post chain
cap rbgs_L12 12 rbgs f64 pre post
python tools/profile_ops.py 12 jacobi f64 sweep residual norm restrict prolong > $O/r02f_unfused.plain.log 2>&1 && timeout 600 $NCU -k regex:"k_jacobi|k_residual|k_restrict|k_prolong" -c 12 -o $O/r02f_unfused python tools/profile_ops.py 12 jacobi f64 sweep residual norm restrict prolong > $O/r02f_unfused.ncu.log 2>&1
python tools/profile_ops.py 12 jacobi f64 cycle > $O/r02f_tail.plain.log 2>&1 && timeout 600 $NCU -k regex:"k_tail" -c 2 -o $O/r02f_tail python tools/profile_ops.py 12 jacobi f64 cycle > $O/r02f_tail.ncu.log 2>&1
# launch list of V-cycles (default bench command), times only
CMD="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3"
$CMD > $O/r02f_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 900 --csv --log-file $O/r02f_launches.csv $CMD > $O/r02f_ncu_launches.log 2>&1
ls -la $O | grep r02f | head -30
