"""Two independent transcriptions of etcd-raft must agree: the C oracle (oracle/raft_oracle.c: arrays, run-length
log) against tests/pyraft_model.py (maps, per-entry log, upstream's own structure), on the synthetic traces and on
adversarial random message soups (every type, stale / equal / higher terms, rejects, out-of-range indices)."""
import numpy as np
import pytest

import oracle
from oracle import Oracle, TraceParams
from pyraft_model import PyEngine

COLUMNS = ("term", "vote", "committed", "last_index", "last_term", "role", "lead", "election_elapsed", "heartbeat_elapsed",
           "randomized_timeout", "out")


def preset(cfg):
    p = TraceParams()
    p.seed = 0x5EED0000 + cfg
    p.p_ack_256, p.p_grant_256, p.p_reject_256, p.p_heartbeat_256 = 256, 230, 0, 0
    p.churn_65536, p.lagging_pct, p.max_prop, p.lag_kind = 0, 0, 3, 0
    if cfg == 5:
        p.p_grant_256, p.p_reject_256, p.churn_65536, p.lagging_pct, p.lag_kind = 205, 26, 400, 20, 1
    if cfg == 6:  # follower heavy: deposed by heartbeats, then heart-beaten
        p.p_grant_256, p.p_reject_256, p.churn_65536, p.p_heartbeat_256, p.lag_kind = 150, 60, 900, 235, 1
    return p


def compare(o, m, where):
    s = o.export()
    for k in COLUMNS:
        got = np.asarray(m.column(k), dtype=np.uint64)
        want = s[k].astype(np.uint64)
        assert np.array_equal(got, want), f"{where}: {k}: model {got[got != want][:4]} oracle {want[got != want][:4]}"
    for g, r in enumerate(m.groups):
        if r.state == 2:
            assert [r.prs[p].Match for p in range(1, m.R + 1)] == [int(x) for x in s["match"][:, g]], f"{where}: match g={g}"
        votes = [0 if p not in r.votes else (1 if r.votes[p] else 2) for p in range(1, m.R + 1)]
        assert votes == [int(x) for x in s["votes"][:, g]], f"{where}: votes g={g}"
    assert sum(r.errors for r in m.groups) == o.errors


@pytest.mark.parametrize("R,cfg", [(1, 2), (2, 5), (3, 2), (3, 6), (4, 5), (5, 5), (5, 6), (7, 5), (8, 5)])
def test_synthetic_traces(R, cfg):
    G, T = 40, 260
    o = Oracle(G, R, seed=77 + R, group_base=1000)
    m = PyEngine(G, R, seed=77 + R, group_base=1000)
    p = preset(cfg)
    for t in range(T):
        ib = o.gen_trace(p, t)
        o.tick(ib)
        m.tick(ib)
        compare(o, m, f"R={R} cfg={cfg} tick {t}")
    assert any(r.state == 2 for r in m.groups)


@pytest.mark.parametrize("R,seed", [(3, 1), (5, 2), (7, 3), (4, 4)])
def test_random_message_soup(R, seed):
    """Nothing protocol-shaped about it: any type from any sender with terms around the receiver's, rejects,
    indices below / at / beyond the log, bursts of proposals.  Both transcriptions must still agree."""
    rng = np.random.default_rng(seed)
    G, T = 48, 220
    o = Oracle(G, R, seed=seed, election_tick=5)
    m = PyEngine(G, R, seed=seed, election_tick=5)
    types = np.array([0, 0, 0, 3, 4, 4, 4, 5, 6, 6, 6, 6, 6, 8, 9, 4 | 0x80, 6 | 0x80, 3 | 0x80], np.uint8)
    seen = set()
    for t in range(T):
        cur = o.export()
        ib = oracle.empty_inbox(G, R)
        ib["type"][:] = rng.choice(types, size=(R, G))
        # mostly the receiver's own term; now and then stale or ahead (a storm of higher terms would keep
        # every group a follower for ever and never reach the candidate / leader code)
        dt = rng.choice(np.array([-2, -1] + [0] * 60 + [1, 2], np.int64), size=(R, G))
        ib["term"][:] = np.maximum(cur["term"][None, :].astype(np.int64) + dt, 0).astype(np.uint64)
        di = rng.integers(-3, 4, size=(R, G))
        ib["index"][:] = np.maximum(cur["last_index"][None, :].astype(np.int64) + di, 0).astype(np.uint64)
        ib["logterm"][:] = np.maximum(cur["last_term"][None, :].astype(np.int64) + rng.integers(-1, 2, size=(R, G)), 0).astype(np.uint64)
        # commits at or a little beyond the log (beyond = upstream would panic: both count an error and skip)
        ib["commit"][:] = np.maximum(cur["committed"][None, :].astype(np.int64) + rng.integers(0, 4, size=(R, G)), 0).astype(np.uint64)
        # a host-resolved MsgApp carries a commit no larger than its new last index
        app = (ib["type"] & 0x0F) == 3
        ib["commit"][app] = np.minimum(ib["commit"][app], ib["index"][app])
        ib["prop_count"][:] = rng.choice(np.array([0, 0, 1, 4], np.uint32), size=G)
        absent = (ib["type"] & 0x0F) == 0
        for k in ("term", "index", "logterm", "commit"):
            ib[k][absent] = 0
        o.tick(ib)
        m.tick(ib)
        compare(o, m, f"soup R={R} tick {t}")
        seen.update(r.state for r in m.groups)
    assert seen == {0, 1, 2}  # the soup really drives groups through all three roles
