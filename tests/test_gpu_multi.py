"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): sharded engines + the per-tick all-gather of
committed[] must equal one engine running every group, bit for bit, for both gather implementations."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, G_total, R, cfg, T, seed, mode, q, tick_mode=0):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from raftsql_b200 import Engine, multi, preset_trace

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        base, G = multi.shard_range(G_total, rank, world)
        eng = Engine(G, R, seed=seed, group_base=base, device=rank)
        multi.attach(eng, dist, mode)
        p = preset_trace(cfg)
        if tick_mode == 4:  # compact state + byte frames: the fused gather rides the quad kernel's 16-byte peer stores
            from raftsql_b200.packed import Pack8

            eng.set_tick_mode(4)
            pk = None
            for t in range(T):
                eng.gen_trace(p, t)
                ib = eng.read_inbox()
                if t % 25 == 0:
                    s = eng.export_state()
                    base = np.where(s["last_index"] > 30, s["last_index"] - np.uint64(30), 0).astype(np.uint64)
                    pk = Pack8(s["self_id"], base, s["term"], R)
                    eng.set_packed_base(pk.base_index, pk.base_term)
                word, prop8, wide = pk.frame(ib)
                eng.post_inbox_packed(word, prop8, wide, slot=1)
                eng.tick(1)
        else:
            for t in range(T):
                eng.gen_trace(p, t)
                eng.tick()
        eng.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        gathered = eng.sync_gathered()
        ref = multi.gather_reference(dist, eng.sync_commits())
        ok = bool(np.array_equal(gathered, ref))
        oks = [None] * world
        dist.all_gather_object(oks, ok)
        if rank == 0:
            q.put((gathered, oks))
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,tick_mode", [("nccl", 0), ("fused", 0), ("fused", 4), ("nccl", 4)])
def test_sharded_engines_gather_equals_single_engine(mode, tick_mode):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp

    from raftsql_b200 import Engine, preset_trace

    world = min(_ngpu(), 8)
    world = 8 if world >= 8 else (4 if world >= 4 else 2)
    G_total, R, cfg, T, seed = 1 << 16, 5, 5, 150, 0xC0FFEE
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, G_total, R, cfg, T, seed, mode, q, tick_mode)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, oks = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(oks), f"gather differs from the host-side concatenation on some rank: {oks}"
    with Engine(G_total, R, seed=seed) as eng:
        p = preset_trace(cfg)
        for t in range(T):
            eng.gen_trace(p, t)
            eng.tick()
        want = eng.sync_commits()
    np.testing.assert_array_equal(gathered, want)
    assert (want > 0).mean() > 0.5
