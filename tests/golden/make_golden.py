#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

What they are (and are not): the reference (a Go adapter around un-vendored etcd-raft) holds NO golden
vector, known-answer test or fixture for this path — its two tests assert SQL text only
(raftsql_test.go:109-111,144,155,167) — and it cannot be run here (no Go toolchain).  So these fixtures are
(1) `upstream_kats.json`: etcd-raft's own test tables for the path, recalled in SURVEY §8c and re-derived
    by hand (the primary pin of the oracle, also replayed straight through the GPU engine); and
(2) `trace_*.npz`: regression vectors produced by the CPU oracle (oracle/raft_oracle.c) on small seeded
    traces — per-tick digests plus the final state — so that the oracle cannot drift silently and the GPU
    engine can be checked against committed bytes as well as against a live oracle run.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import Oracle, TraceParams  # noqa: E402

DIGEST_COLUMNS = ("term", "vote", "committed", "last_index", "last_term", "term_start", "role", "lead", "votes",
                  "election_elapsed", "heartbeat_elapsed", "randomized_timeout", "out")

CASES = [  # name, G, R, preset, ticks, seed
    ("trace_512x3_cfg2", 512, 3, 2, 240, 0x5EED0002),
    ("trace_512x5_cfg5", 512, 5, 5, 320, 0x5EED0005),
    ("trace_384x7_cfg5", 384, 7, 5, 320, 0x5EED0005),
    ("trace_300x4_cfg5", 300, 4, 5, 200, 0x5EED0004),
]


def preset(cfg):
    p = TraceParams()
    p.seed = 0x5EED0000 + cfg
    p.p_ack_256, p.p_grant_256, p.p_reject_256, p.p_heartbeat_256 = 256, 230, 0, 0
    p.churn_65536, p.lagging_pct, p.max_prop, p.lag_kind = 0, 0, 3, 0
    if cfg in (3, 4):
        p.lag_kind = 1
    elif cfg == 5:
        p.p_grant_256, p.p_reject_256, p.churn_65536, p.lagging_pct, p.lag_kind = 205, 26, 43, 20, 1
    return p


def digest(state: dict) -> str:
    h = hashlib.sha256()
    for k in DIGEST_COLUMNS:
        h.update(np.ascontiguousarray(state[k]).tobytes())
    lead = state["role"] == 2  # Progress.Match is defined for leaders only
    h.update(np.ascontiguousarray(state["match"][:, lead]).tobytes())
    return h.hexdigest()


def run_case(name, G, R, cfg, T, seed):
    o = Oracle(G, R, seed=seed)
    p = preset(cfg)
    digests = []
    for t in range(T):
        o.tick(o.gen_trace(p, t))
        digests.append(digest(o.export()))
    s = o.export()
    assert o.errors == 0
    np.savez_compressed(os.path.join(HERE, name + ".npz"), digests=np.array(digests),
                        meta=np.array([G, R, cfg, T, seed], dtype=np.uint64), **{"final_" + k: v for k, v in s.items()})
    print(f"{name}: leaders {(s['role'] == 2).mean():.2f}, max term {s['term'].max()}, digest[-1] {digests[-1][:16]}")


UPSTREAM_KATS = {
    "_source": "etcd raft/raft_test.go and raft/raft_paper_test.go (v2.2-v2.3 era), recalled in SURVEY.md 8c "
               "[UPSTREAM-RECALLED] and re-derived by hand from SURVEY 8a rows a7-a16; the upstream source is "
               "not available in this environment",
    "TestCommit": [  # matches, log entry terms, smTerm, want committed
        [[1], [1], 1, 1], [[1], [1], 2, 0], [[2], [1, 2], 2, 2], [[1], [2], 2, 1],
        [[2, 1, 1], [1, 2], 1, 1], [[2, 1, 1], [1, 1], 2, 0], [[2, 1, 2], [1, 2], 2, 2], [[2, 1, 2], [1, 1], 2, 0],
        [[2, 1, 1, 1], [1, 2], 1, 1], [[2, 1, 1, 1], [1, 1], 2, 0], [[2, 1, 1, 2], [1, 2], 1, 1],
        [[2, 1, 1, 2], [1, 1], 2, 0], [[2, 1, 2, 2], [1, 2], 2, 2], [[2, 1, 2, 2], [1, 1], 2, 0]],
    "TestVoter": [  # voter log terms, candidate logterm, candidate index, want reject
        [[1], 1, 1, False], [[1], 1, 2, False], [[1, 1], 1, 1, True], [[1], 2, 1, False], [[1], 2, 2, False],
        [[1, 1], 2, 1, False], [[2], 1, 1, True], [[2], 1, 2, True], [[2, 1], 1, 1, True]],
    "TestFollowerVote": [  # vote, nvote, want reject
        [0, 1, False], [0, 2, False], [1, 1, False], [2, 2, False], [1, 2, True], [2, 1, True]],
    "TestLeaderElectionInOneRoundRPC": [  # cluster size, {voter id: granted}, want state (0 F, 1 C, 2 L)
        [1, {}, 2], [3, {"2": True, "3": True}, 2], [3, {"2": True}, 2],
        [5, {"2": True, "3": True, "4": True, "5": True}, 2], [5, {"2": True, "3": True, "4": True}, 2],
        [5, {"2": True, "3": True}, 2], [3, {"2": False, "3": False}, 0],
        [5, {"2": False, "3": False, "4": False, "5": False}, 0],
        [5, {"2": True, "3": False, "4": False, "5": False}, 0], [3, {}, 1], [5, {"2": True}, 1],
        [5, {"2": False, "3": False}, 1], [5, {}, 1]],
}


if __name__ == "__main__":
    with open(os.path.join(HERE, "upstream_kats.json"), "w") as f:
        json.dump(UPSTREAM_KATS, f, indent=1)
    for c in CASES:
        run_case(*c)
