"""torch.distributed plumbing for the one-process-per-GPU row-slab mode.

torch.distributed is used only to bootstrap (broadcast the 128-byte communicator id) and,
in tests/bench, for barriers and max-reductions of timings.  All data-path communication
(halo rows, coarse all-gather, norm) happens inside libmgb200.so over NCCL on the
context's own stream (csrc/comm.cu)."""
from __future__ import annotations

import os

import numpy as np

from .solver import Multigrid, comm_id


def env_ranks():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def create(finest_level: int, **kw) -> Multigrid:
    """Collectively create one context per rank (call on every rank of an initialised
    torch.distributed process group; with world size 1 it is a plain Multigrid)."""
    import torch.distributed as dist
    rank, world, local = env_ranks()
    if world == 1 or not dist.is_initialized():
        return Multigrid(finest_level, **kw)
    blob = [comm_id() if rank == 0 else None]
    dist.broadcast_object_list(blob, src=0)
    kw.setdefault("device", local)
    return Multigrid(finest_level, rank=rank, world=world, comm_id=blob[0], **kw)


def owned_rows(mg: Multigrid, level: int):
    """1-based interior rows [a, b) this rank owns on `level`."""
    from . import capi
    return mg.info(capi.MG_INFO_ROW_BEGIN, level), mg.info(capi.MG_INFO_ROW_END, level)


def owned_slice(mg: Multigrid, level: int, vec: np.ndarray) -> np.ndarray:
    """The rows of a full interior vector that this rank owns (a view)."""
    n = (1 << level) - 1
    a, b = owned_rows(mg, level)
    return vec.reshape(n, n)[a - 1:b - 1]
