"""GPU parity of the multi-GPU schedule variants that are selected by environment variables, under torchrun via
tests/mgpu_worker.py (every rank's rows must equal the single-domain oracle's, bit for bit)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"MGB200_COMM_AVOID": "0"}, {"MGB200_GRAPH_DIST": "0"},
                                 {"MGB200_COMM_AVOID": "0", "MGB200_GRAPH_DIST": "0"}, {"MGB200_CHAIN": "0", "MGB200_ZERO_GUESS": "0"}])
def test_multigpu_optin_paths(env):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env={**os.environ, **env})
    assert out.returncode == 0 and "MGPU OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
