"""CPU tests of the N>1 path (gloo, world size 2 and 4): the row-slab schedule -- partition
from the C ABI (mg_slab_rows), one halo row exchanged after every sweep / colour, the
neighbour's edge residual row before restriction, no exchange after prolongation,
all-gather at the agglomeration level -- reproduces the single-domain oracle bit for bit
from NaN-poisoned slabs.  (The CUDA kernels themselves need a GPU: tests/test_multigpu.py.)"""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    """a TCP port nobody listens on right now (rendezvous of the gloo process group)"""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, level, aggl, smoother, gamma, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mgb200
        import oracle
        from conftest import rand_vec
        from slab_emulator import SlabEmu
        o = oracle.get()
        p = oracle.Params(smoother=smoother, gamma=gamma)
        x, b = rand_vec(level, np.float64, 51), rand_vec(level, np.float64, 52, 1e-3)
        emu = SlabEmu(mgb200.capi.lib(), o, level, aggl, rank, world, p)
        emu.set("u", level, x)
        emu.set("f", level, b)
        want = x
        for k in range(2):
            emu.cycle(level)
            want = o.vcyclemultigrid(want, b, p)
            got = emu.owned(level, emu.u[level])
            assert not np.isnan(got).any(), "owned rows depend on data outside the slab"
            assert np.array_equal(got, emu.owned(level, want)), f"cycle {k}"
        emu.set("f", level, b)
        emu.fmg(1)
        assert np.array_equal(emu.owned(level, emu.u[level]), emu.owned(level, o.fullmultigrid(b, 1, p))), "fmg"
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"FAIL: {type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,level,aggl,smoother,gamma", [
    (2, 6, 3, 0, 1), (2, 6, 4, 1, 1), (2, 7, 5, 0, 2), (4, 7, 4, 0, 1), (4, 7, 5, 1, 2),
])
def test_slab_schedule_matches_single_domain(world, level, aggl, smoother, gamma):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, level, aggl, smoother, gamma, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert all(msg == "ok" for _, msg in res), res
