"""Generate tests/golden/ref_pins.json (+ ref_interp.npz) from the REFERENCE'S OWN SOURCE
compiled against the stub oneMKL/SYCL headers (oracle/_ref/libref_poisson.so; `make -C oracle
ref`, needs /root/reference).  Run from the repo root:  python tests/golden/make_ref_golden.py
The full-program run (levels 7..10, mu0=30, mu1=mu2=10, fp32, as written) takes ~30 s."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_poisson.so"))
for fn in ("ref_globalforcefunction", "ref_level_csr_stats", "ref_run_program"):
    getattr(R, fn).restype = ctypes.c_longlong


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


pins = {"source": "/root/reference/Poissons_SYCL.cpp compiled with oracle/stub (g++ -O2 -ffp-contract=off)",
        "finest_level": R.ref_finest_level(), "coarsest_level": R.ref_coarsest_level()}
par = (ctypes.c_int * 3)()
R.ref_params(par)
pins["mu0_mu1_mu2"] = list(par)

# interpolation2d (P:337-425): outputs on seeded inputs
arrays = {}
for m in (1, 2, 3, 7, 15, 31):
    x = np.random.default_rng(1000 + m).uniform(-1, 1, m * m).astype(np.float32)
    out = np.zeros((2 * m + 1) ** 2, np.float32)
    R.ref_interpolation2d(P(x), m, P(out))
    arrays[f"interp_in_{m}"] = x
    arrays[f"interp_out_{m}"] = out
# restriction2d as written (P:539: (1/16) == 0)
x = np.random.default_rng(7).uniform(-1, 1, 31 * 31).astype(np.float32)
out = np.ones(15 * 15, np.float32)
R.ref_restriction2d(P(x), 31, P(out))
pins["restriction2d_as_written_max_abs"] = float(np.abs(out).max())
# globalforcefunction (P:283-335) at the reference's finest level
n = (1 << pins["finest_level"]) - 1
f = np.zeros(n * n, np.float32)
pins["globalforce_size"] = int(R.ref_globalforcefunction(P(f)))
pins["globalforce_min"], pins["globalforce_max"] = float(f.min()), float(f.max())
# assembled operators (P:200-281 + P:55-116): E1 / E3 facts
pins["csr"] = {}
for lvl in (2, 3, 4):
    for which, name in ((0, "lu"), (1, "d")):
        st = (ctypes.c_double * 4)()
        rows, mr = ctypes.c_int(), ctypes.c_int()
        nnz = R.ref_level_csr_stats(lvl, which, st, ctypes.byref(rows), ctypes.byref(mr))
        pins["csr"][f"L{lvl}_{name}"] = {"nnz": int(nnz), "min": st[0], "max": st[1], "sum": st[2], "coo_sum": st[3],
                                         "rows": rows.value, "max_row_nnz": mr.value}
# call structure of one vcyclemultigrid (P:575-627) at level 8 and 9
pins["vcycle_calls"] = {}
for lvl in (8, 9):
    nn = (1 << lvl) - 1
    u = np.zeros(nn * nn, np.float32)
    b = np.full(nn * nn, np.float32(-4.0 / 4 ** lvl))
    R.ref_counters_reset()
    R.ref_vcyclemultigrid(lvl, P(u), P(b))
    c = (ctypes.c_longlong * 5)()
    R.ref_counters(c)
    pins["vcycle_calls"][f"L{lvl}"] = {"gemv": c[0], "scal": c[1], "add": c[2], "sub": c[3], "u_min": float(u.min()),
                                       "u_max": float(u.max())}
# jacobirelaxation (P:125-147), the reference's own body, on the INTENDED off-diagonal operator (A_lu = -1 per neighbour)
def intended_a_lu(m):
    rp, ci = [0], []
    for r in range(m):
        for c in range(m):
            for rr, cc in ((r - 1, c), (r, c - 1), (r, c + 1), (r + 1, c)):
                if 0 <= rr < m and 0 <= cc < m:
                    ci.append(rr * m + cc)
            rp.append(len(ci))
    return np.array(rp, np.int32), np.array(ci, np.int32), np.full(len(ci), -1.0, np.float32)


for m, mu in ((7, 1), (15, 3), (31, 10)):
    rp, ci, va = intended_a_lu(m)
    v = np.random.default_rng(2000 + m).uniform(-1, 1, m * m).astype(np.float32)
    fh = (1e-2 * np.random.default_rng(3000 + m).uniform(-1, 1, m * m)).astype(np.float32)
    arrays[f"jacobi_v_{m}"], arrays[f"jacobi_f_{m}"] = v.copy(), fh
    R.ref_jacobirelaxation_with(m * m, P(rp), P(ci), P(va), P(v), P(fh), mu)
    arrays[f"jacobi_out_{m}_mu{mu}"] = v
# the whole program as written
t0 = time.time()
sol = np.zeros(n * n, np.float32)
R.ref_counters_reset()
size = R.ref_run_program(P(sol))
c = (ctypes.c_longlong * 5)()
R.ref_counters(c)
pins["program"] = {"size": int(size), "min": float(sol.min()), "max": float(sol.max()), "gemv": c[0], "scal": c[1],
                   "add": c[2], "sub": c[3], "seconds": round(time.time() - t0, 1)}
json.dump(pins, open(os.path.join(ROOT, "tests", "golden", "ref_pins.json"), "w"), indent=1)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_interp.npz"), **arrays)
print(json.dumps(pins, indent=1))
