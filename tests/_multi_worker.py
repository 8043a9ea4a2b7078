"""Worker for tests/test_multi_cpu.py: one process per rank, gloo backend, CPU only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def preset(mod, cfg):
    p = mod.TraceParams()
    p.seed = 0x5EED0000 + cfg
    p.p_ack_256, p.p_grant_256, p.p_reject_256, p.p_heartbeat_256 = 256, 230, 0, 0
    p.churn_65536, p.lagging_pct, p.max_prop, p.lag_kind = 0, 0, 3, 0
    if cfg == 5:
        p.p_grant_256, p.p_reject_256, p.churn_65536, p.lagging_pct, p.lag_kind = 205, 26, 43, 20, 1
    return p


def worker(rank, world, port, G_total, R, cfg, T, seed, q):
    import torch.distributed as dist

    import oracle
    from raftsql_b200 import multi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base, G = multi.shard_range(G_total, rank, world)
        o = oracle.Oracle(G, R, seed=seed, group_base=base)
        p = preset(oracle, cfg)
        for t in range(T):
            o.tick(o.gen_trace(p, t))
        s = o.export()
        gathered = multi.gather_reference(dist, s["committed"])
        terms = multi.gather_reference(dist, s["term"])
        # out-of-band plumbing used for the NCCL id / IPC handles
        blobs = multi.exchange_bytes(dist, bytes([rank + 1]) * 64, world)
        assert [b[0] for b in blobs] == [r + 1 for r in range(world)]
        uid = multi.broadcast_bytes(dist, bytes(range(128)) if rank == 0 else None, 128, 0)
        assert uid == bytes(range(128))
        if rank == 0:
            q.put((gathered, terms))
        dist.barrier()
    finally:
        dist.destroy_process_group()
