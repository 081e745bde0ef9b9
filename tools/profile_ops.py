"""Run a few named operators of libmgb200 back to back so that ncu can capture exactly them:
    ncu --set full --clock-control none --import-source on -k regex:k_stream -c 12 -o out python tools/profile_ops.py 12 rbgs f64 pre post
Tuner off (MGB200_AUTOTUNE=0 is set here) so that no tuning launches pollute the capture."""
import os
import sys

os.environ.setdefault("MGB200_AUTOTUNE", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import mgb200  # noqa: E402
from mgb200 import capi  # noqa: E402

OPS = {"pre": capi.MG_OP_PRE_FUSED, "post": capi.MG_OP_POST_FUSED, "chain": capi.MG_OP_POSTPRE_FUSED, "sweep": capi.MG_OP_SMOOTH1,
       "sweep2": capi.MG_OP_SMOOTH2, "residual": capi.MG_OP_RESIDUAL, "norm": capi.MG_OP_RESIDUAL_NORM, "restrict": capi.MG_OP_RESTRICT,
       "prolong": capi.MG_OP_PROLONG}
level, smoother, dt = int(sys.argv[1]), sys.argv[2], sys.argv[3]
with mgb200.Multigrid(level, smoother=smoother, dtype=np.float64 if dt == "f64" else np.float32) as mg:
    mg.force_synthetic(1234)
    mg.zero_u(level)
    for name in sys.argv[4:]:
        if name == "cycle":
            mg.cycle(level, 2, 2, 1)
            mg.cycle(level, 2, 2, 1)
        elif name == "wcycle":
            mg.cycle(level, 2, 2, 2)
        else:
            ms = mg.time_op(OPS[name], level, 2)
            print(name, f"{ms / 2 * 1e3:.1f} us per launch")
    mg.sync()
