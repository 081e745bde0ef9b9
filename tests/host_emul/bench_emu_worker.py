"""tests/host_emul/bench_emu_worker.py — TEST INFRASTRUCTURE.  Runs bench.py's `run_ours` / `run_micro` end to end on
CPU ranks against the emulated library (real `Multigrid`, real C ABI, real slab logic; only torch.cuda and the NCCL
bootstrap are replaced), at a small level.  Under torchrun for world > 1.  The numbers are meaningless; what is
checked is that the harness runs and that the JSON line is well formed."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["MGB200_TEST_EMU"] = "1"

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import conftest  # noqa: E402

conftest.use_emulated_library()
import bench  # noqa: E402

torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
_init = dist.init_process_group
dist.init_process_group = lambda backend, **kw: _init("gloo")
_tensor = torch.tensor
torch.tensor = lambda data, **kw: _tensor(data, **{k: v for k, v in kw.items() if k != "device"})
bench.pinned = lambda nelem, dtype, device=None: (None, np.empty(nelem, dtype=dtype))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=7)
    ap.add_argument("--micro", action="store_true")
    ap.add_argument("--smoother", default="jacobi")
    ap.add_argument("--gamma", type=int, default=1)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--aggl", type=int, default=0)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    args = argparse.Namespace(gpus=world, steps=3, warmup=1, impl="ours", level=a.level, dtype=a.dtype, smoother=a.smoother,
                              nu1=2, nu2=2, gamma=a.gamma, no_graph=False, no_fused=False, no_tail=False, no_cpu=True,
                              aggl=a.aggl, no_e2e=False, full_host_vectors=False, no_n1=False, no_extra=False, no_phases=False, micro=a.micro, leg_timeout=120.0)
    if a.micro:
        bench.run_micro(args, rank, world, 0)
    else:
        bench.run_ours(args, rank, world, 0)


if __name__ == "__main__":
    main()
