#!/bin/bash
# 1-GPU call: ncu evidence for the shipped kernels, per kernel and per level (each command first runs clean without ncu)
set -u
mkdir -p gpurun_out; O=gpurun_out
# programmatic dependent launches: A/B on the cycle time (isolated and chained), V and W cycles
python - <<'PY'
import os, sys, statistics
sys.path.insert(0, '.')
import mgb200
for pdl in ("0", "1"):
    os.environ["MGB200_PDL"] = pdl
    out = []
    for level, gamma, sm in ((6, 1, "jacobi"), (9, 1, "jacobi"), (12, 1, "jacobi"), (12, 1, "rbgs"), (13, 2, "jacobi")):
        with mgb200.Multigrid(level, smoother=sm) as mg:
            mg.force_synthetic(1234); mg.zero_u(level)
            mg.time_cycle(level, 2, 2, gamma, 3)
            iso = statistics.median([mg.time_cycle(level, 2, 2, gamma, 1) for _ in range(15)])
            mg.time_cycle(level, 2, 2, gamma, 10)
            ch = statistics.median([mg.time_cycle(level, 2, 2, gamma, 10) for _ in range(7)]) / 10
            c0 = mg.checksum(level, 0)
            out.append(f"L{level}g{gamma}{sm[0]}: iso {iso*1e3:.1f} chained {ch*1e3:.1f} us csum {c0:016x}")
    print(f"PDL={pdl}", " | ".join(out))
PY
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
