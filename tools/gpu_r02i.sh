#!/bin/bash
# resident streaming warps per SM: 12 (default) vs 13 / 14 / 16 (registers allow 14 for the Jacobi chain kernel at 142)
set -u
mkdir -p gpurun_out
for occ in 12 14 16; do
MGB200_STREAM_OCC=$occ python - $occ <<'PY'
import sys, statistics
sys.path.insert(0, '.')
import mgb200
from mgb200 import capi
out = []
for level, sm in ((12, "jacobi"), (12, "rbgs")):
    with mgb200.Multigrid(level, smoother=sm) as mg:
        mg.force_synthetic(1234); mg.zero_u(level)
        mg.time_cycle(level, 2, 2, 1, 3)
        iso = statistics.median([mg.time_cycle(level, 2, 2, 1, 1) for _ in range(11)])
        mg.time_cycle(level, 2, 2, 1, 10)
        ch = statistics.median([mg.time_cycle(level, 2, 2, 1, 10) for _ in range(7)]) / 10
        ops = []
        for name, op in (("pre", capi.MG_OP_PRE_FUSED), ("post", capi.MG_OP_POST_FUSED), ("chain", capi.MG_OP_POSTPRE_FUSED)):
            try:
                mg.time_op(op, level, 2)
                ops.append(f"{name} {mg.time_op(op, level, 10) / 10 * 1e3:.0f}")
            except Exception:
                pass
        out.append(f"L{level}{sm[0]}: iso {iso*1e3:.1f} chained {ch*1e3:.1f} [{' '.join(ops)}]")
print(f"OCC={sys.argv[1]}", " | ".join(out), flush=True)
PY
done
