"""Time the streaming kernels of one level for the current MGB200_STREAM_* environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mgb200
from mgb200 import capi
level = int(sys.argv[1]) if len(sys.argv) > 1 else 12
mg = mgb200.Multigrid(level)
n = (1 << level) - 1
mg.force_constant(4.0)
mg.zero_u(level)
mg.cycle(level)
out = []
for name, op in (("sw2", capi.MG_OP_SMOOTH2), ("pre", capi.MG_OP_PRE_FUSED), ("post", capi.MG_OP_POST_FUSED)):
    mg.time_op(op, level, 5)
    out.append(f"{name} {mg.time_op(op, level, 20) / 20 * 1000:6.1f}")
mg.time_cycle(level, 2, 2, 1, 5)
out.append(f"cycle {mg.time_cycle(level, 2, 2, 1, 20) / 20 * 1000:6.1f}")
print(f"RY={os.environ.get('MGB200_STREAM_RY', '-'):>4} OCC={os.environ.get('MGB200_STREAM_OCC', '-'):>3} L{level}: " + "  ".join(out), flush=True)
