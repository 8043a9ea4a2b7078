"""Multi-GPU host plumbing: one process per GPU, groups sharded contiguously, one all-gather of the
committed indices per tick.

The data path lives in libmrq.so (ncclAllGather on the engine's stream, or peer stores fused into the tick
kernel over CUDA-IPC mapped buffers).  This module only does what a host has to do around it: the shard
arithmetic and the out-of-band exchange of the NCCL id / IPC handles, over whatever `torch.distributed`
backend the launcher initialised (nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_groups_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard of rank `rank`: (group_base, n_groups).  Ranks differ by at most one group."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = n_groups_total * rank // world
    hi = n_groups_total * (rank + 1) // world
    return lo, hi - lo


def exchange_bytes(dist, payload: bytes, world: int) -> list[bytes]:
    """All-gather a small fixed-size byte string across ranks (works on gloo and nccl)."""
    import torch

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def broadcast_bytes(dist, payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    import torch

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if payload is not None:
        t.copy_(torch.tensor(list(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def attach(engine, dist, mode: str = "fused") -> None:
    """Attach `engine` (this rank's shard) to the job-wide gather of committed[].

    mode "nccl":  mrq_comm_init -> every mrq_tick ends with ncclAllGather(committed) on the engine stream.
    mode "fused": CUDA-IPC handles of every rank's gather buffer are exchanged and mapped; the tick kernel
                  then stores each commit index straight into all ranks' buffers over NVLink.
    Every rank must own the same number of groups (the gather is a dense [world][G] vector)."""
    from . import _ffi
    from .engine import Engine

    world, rank = dist.get_world_size(), dist.get_rank()
    if mode == "nccl":
        uid = Engine.comm_unique_id() if rank == 0 else None
        uid = broadcast_bytes(dist, uid, _ffi.MRQ_COMM_ID_BYTES, 0)
        engine.comm_init(uid, rank, world)
    elif mode == "fused":
        engine.ipc_prepare(world)
        handles = exchange_bytes(dist, engine.ipc_export(), world)
        engine.ipc_attach(b"".join(handles), rank, world)
        engine.comm_set_mode(1)
    else:
        raise ValueError(mode)


def gather_reference(dist, committed: np.ndarray) -> np.ndarray:
    """What the in-engine gather must equal: the concatenation of every rank's committed[] (host path)."""
    import torch

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.from_numpy(committed.view(np.int64)).to(dev)
    outs = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, mine)
    return torch.cat(outs).cpu().numpy().view(np.uint64)
