"""Freeze oracle outputs on seeded inputs into tests/golden/oracle_golden.npz.
Run from the repo root:  python tests/golden/make_golden.py
(The reference has no golden vectors of its own; these fixtures pin the
oracle against drift and give the GPU parity tests committed expected values.)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from conftest import rand_vec  # noqa: E402

o = oracle.get()
cases, arrays = [], {}
k = 0
for dtype in ("float64", "float32"):
    for level in (3, 5, 6):
        for op, smoother, gamma in (("jacobi3", 0, 1), ("rbgs2", 1, 1), ("residual", 0, 1), ("restrict", 0, 1),
                                    ("prolong", 0, 1), ("vcycle", 0, 1), ("vcycle", 1, 1), ("vcycle", 0, 2),
                                    ("fmg", 0, 1)):
            k += 1
            case = dict(name=f"c{k:03d}_{op}_L{level}_{dtype}_s{smoother}_g{gamma}", op=op, level=level, dtype=dtype,
                        seed_u=100 + k, seed_b=200 + k, scale_b=1e-3, smoother=smoother, gamma=gamma)
            x = rand_vec(level, np.dtype(dtype), case["seed_u"])
            b = rand_vec(level, np.dtype(dtype), case["seed_b"], case["scale_b"])
            p = oracle.Params(smoother=smoother, gamma=gamma)
            arrays[case["name"]] = {"jacobi3": lambda: o.jacobirelaxation(x, b, 3),
                                    "rbgs2": lambda: o.rbgs(x, b, 2),
                                    "residual": lambda: o.residual(x, b),
                                    "restrict": lambda: o.restriction2d(x),
                                    "prolong": lambda: o.interpolation2d(x),
                                    "vcycle": lambda: o.vcyclemultigrid(x, b, p),
                                    "fmg": lambda: o.fullmultigrid(b, 1, p)}[op]()
            cases.append(case)
arrays["meta"] = np.array(json.dumps({"cases": cases}))
out = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
np.savez_compressed(out, **arrays)
print("wrote", out, os.path.getsize(out), "bytes,", len(cases), "cases")
