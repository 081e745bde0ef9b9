"""Row-slab multi-GPU parity: results on 2/4/8 GPUs are bit-identical to the
oracle (hence to the single-GPU run) for any agglomeration level.  Needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_match_oracle(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:   # a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MGPU OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_two_contexts_on_two_devices_in_one_process():
    """Per-context (per-device) kernel attributes and tuning state: a second context on another device of the same process
    launches the > 48 KB dynamic-shared-memory kernels, and both contexts give the oracle's bits."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    import mgb200
    import oracle
    from conftest import assert_bitwise, rand_vec
    o = oracle.get()
    level = 9
    x, b = rand_vec(level, np.float64, 71), rand_vec(level, np.float64, 72, 1e-3)
    want = o.vcyclemultigrid(x, b, oracle.Params(nthreads=4))
    ctxs = [mgb200.Multigrid(level, device=d) for d in (0, 1)]
    try:
        for rep in range(2):                       # interleaved use of the two contexts
            for mg in ctxs:
                mg.set_u(level, x)
                mg.set_rhs(level, b)
                mg.cycle(level, 2, 2, 1)
            for d, mg in enumerate(ctxs):
                assert_bitwise(mg.get_u(level), want, f"device {d}, round {rep}")
    finally:
        for mg in ctxs:
            mg.close()
