#!/bin/bash
# quick 1-GPU check of what has not been on a GPU yet: exact coarsest solve, staged pageable copies (levels >= 10), device loop default
set -u
mkdir -p gpurun_out; O=gpurun_out
timeout 500 python -m pytest tests/test_parity_gpu.py tests/test_problem_setup.py -m gpu -x -q -k "exact_coarsest or v2_shape or 4097 or (iterates_bitwise and 10-) or (transfer and 10-) or device_side or solve_257 or synthetic or cpp_" --durations=5 > $O/r02h_pytest.log 2>&1; echo "rc=$?" >> $O/r02h_pytest.log; tail -14 $O/r02h_pytest.log
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
import mgb200, torch
level = 12; n = (1 << level) - 1
with mgb200.Multigrid(level) as mg:
    f = (np.random.default_rng(0).uniform(-1, 1, n * n) / (1 << level) ** 2)
    u = np.zeros(n * n)
    ft = torch.empty(n * n, dtype=torch.float64, pin_memory=True); ft.numpy()[:] = f
    ut = torch.zeros(n * n, dtype=torch.float64, pin_memory=True)
    for name, (uu, ff) in (("pageable", (u, f)), ("pinned", (ut.numpy(), ft.numpy()))):
        for _ in range(2): mg.vcyclemultigrid(uu, ff, 2, 2, 1, inplace=True)
        t0 = time.perf_counter()
        for _ in range(5): mg.vcyclemultigrid(uu, ff, 2, 2, 1, inplace=True)
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print(f"e2e mg_host_vcyclemultigrid 4097^2 {name}: {ms:.2f} ms ({402.5 / ms:.1f} GB/s over PCIe)")
    assert np.array_equal(u, ut.numpy())
PY
