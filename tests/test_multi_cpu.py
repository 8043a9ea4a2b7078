"""The N>1 host path on CPU: world_size-2 (and 4) gloo runs of the sharding + gather plumbing.

Groups are independent, so a shard keyed by GLOBAL group ids must reproduce the 1-rank run exactly, and the
gathered commit vector must be the concatenation of the shards' (SURVEY §8e correctness check).  The shards
here are oracle shards (no GPU in this container); the same `multi` helpers attach real engines in bench.py.
"""
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import oracle
from _multi_worker import preset, worker
from raftsql_b200 import multi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_exactly():
    for G in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [multi.shard_range(G, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == G
            for (b0, n0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + n0 == b1
    assert multi.shard_range(1 << 20, 3, 8) == (3 << 17, 1 << 17)
    with pytest.raises(ValueError):
        multi.shard_range(10, 2, 2)


@pytest.mark.parametrize("world,R,cfg", [(2, 5, 5), (2, 3, 2), (4, 7, 5)])
def test_sharded_run_equals_single_run(world, R, cfg):
    G_total, T, seed = 2048, 120, 0xBEEF
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, G_total, R, cfg, T, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, terms = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    o = oracle.Oracle(G_total, R, seed=seed)
    pr = preset(oracle, cfg)
    for t in range(T):
        o.tick(o.gen_trace(pr, t))
    s = o.export()
    np.testing.assert_array_equal(gathered, s["committed"])
    np.testing.assert_array_equal(terms, s["term"])
    assert (s["committed"] > 0).mean() > 0.5
