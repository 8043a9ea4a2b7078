"""Independent-definition property tests for the CPU oracle (SURVEY §8c (i)–(vii)).

These need no upstream source: they pin the *mathematics* of the path — q-th largest == "largest index a
quorum has reached", commit monotone and term-gated, election win/lose exclusive, one vote per term, term
monotone — and that the batched tick equals feeding the same messages one at a time.
"""
import numpy as np
import pytest

import oracle
from oracle import Oracle, TraceParams

FOLLOWER, CANDIDATE, LEADER = 0, 1, 2
U64MAX = np.uint64(0xFFFFFFFFFFFFFFFF)


def _q(R):
    return R // 2 + 1


def _numpy_quorum_index(match_rg):
    """max{x : |{r : match[r] >= x}| >= q}  ==  the q-th largest, via a numpy sort (third definition)."""
    R = match_rg.shape[0]
    return np.sort(match_rg, axis=0)[R - _q(R)]


def test_quorum_index_matches_independent_definition():
    rng = np.random.default_rng(7)
    for R in range(1, 9):
        for _ in range(400):
            kind = rng.integers(0, 4)
            if kind == 0:
                m = rng.integers(0, 4, size=R, dtype=np.uint64)  # many ties
            elif kind == 1:
                m = rng.integers(0, 2 ** 63, size=R, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
            elif kind == 2:
                m = rng.choice(np.array([0, 1, U64MAX, U64MAX - np.uint64(1)], dtype=np.uint64), size=R)
            else:
                m = rng.integers(2 ** 20, 2 ** 40, size=R, dtype=np.uint64)
            brute = oracle.quorum_index_bruteforce(m)
            assert brute == int(_numpy_quorum_index(m.reshape(R, 1))[0])
            # the oracle's sort-descending-and-index form, with a log that makes the gate pass
            got = oracle.kat_commit(m, [1], 1) if int(brute) == 1 else None
            if got is not None:
                assert got == 1


def _leader_state(G, R, rng, term_lo=1, term_hi=8):
    st = oracle.empty_state(G, R)
    g = np.arange(G, dtype=np.uint64)
    st["self_id"][:] = (g % R + 1).astype(np.uint8)
    st["role"][:] = LEADER
    st["lead"][:] = st["self_id"]
    st["term"][:] = rng.integers(term_lo, term_hi + 1, size=G, dtype=np.uint64)
    st["vote"][:] = st["self_id"]
    st["last_index"][:] = rng.integers(2 ** 20, 2 ** 40, size=G, dtype=np.uint64)
    st["last_term"][:] = st["term"]
    lag = rng.geometric(0.2, size=(R, G)).astype(np.uint64)
    st["match"][:] = st["last_index"][None, :] - lag
    st["match"][st["self_id"] - 1, np.arange(G)] = st["last_index"]
    st["committed"][:] = st["last_index"] - np.uint64(40)
    # 99%: the leader's own-term entries start at or below committed; 1%: the gate is still closed
    gate_open = rng.random(G) < 0.99
    st["term_start"][:] = np.where(gate_open, st["committed"] - np.uint64(5), st["last_index"] - np.uint64(1))
    st["randomized_timeout"][:] = 10
    return st


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 7, 8])
def test_quorum_commit_pass_matches_numpy(R):
    G = 3000
    rng = np.random.default_rng(100 + R)
    st = _leader_state(G, R, rng)
    o = Oracle(G, R)
    o.import_state(st)
    o.quorum_commit()
    got = o.export()
    mci = _numpy_quorum_index(st["match"])
    want = np.where((mci > st["committed"]) & (mci >= st["term_start"]), mci, st["committed"])
    np.testing.assert_array_equal(got["committed"], want)
    # (ii) monotone, (iii) gate
    assert (got["committed"] >= st["committed"]).all()
    moved = got["committed"] > st["committed"]
    assert (got["committed"][moved] >= st["term_start"][moved]).all()


def _preset(cfg):
    p = TraceParams()
    p.seed = 0x5EED0000 + cfg
    p.p_ack_256, p.p_grant_256, p.p_reject_256, p.p_heartbeat_256 = 256, 230, 0, 0
    p.churn_65536, p.lagging_pct, p.max_prop, p.lag_kind = 0, 0, 3, 0
    if cfg == 5:
        p.p_grant_256, p.p_reject_256, p.churn_65536, p.lagging_pct, p.lag_kind = 205, 26, 43, 20, 1
    return p


@pytest.mark.parametrize("R,cfg", [(3, 2), (5, 5), (7, 5), (4, 5), (1, 2)])
def test_trace_invariants(R, cfg):
    """Run the synthetic vote/append trace and check (ii)–(vi) on every tick."""
    G, T = 512, 300
    o = Oracle(G, R, seed=0xABCD + R)
    p = _preset(cfg)
    prev = o.export()
    voted = {}  # (g, term) -> vote
    saw_leader = False
    for t in range(T):
        ib = o.gen_trace(p, t)
        o.tick(ib)
        cur = o.export()
        assert (cur["term"] >= prev["term"]).all()  # (vi)
        assert (cur["committed"] >= prev["committed"]).all()  # (ii)
        assert (cur["committed"] <= cur["last_index"]).all()
        lead = cur["role"] == LEADER
        saw_leader |= bool(lead.any())
        moved = cur["committed"] > prev["committed"]
        was_or_is_leader = lead | (prev["role"] == LEADER)
        foll_moved = moved & ~was_or_is_leader
        # (iii) a commit advance by counting replicas only happens on a leader, at or past term_start
        lm = moved & lead
        assert (cur["committed"][lm] >= cur["term_start"][lm]).all()
        # followers only move commit through leader_commit messages, bounded by last_index
        assert (cur["committed"][foll_moved] <= cur["last_index"][foll_moved]).all()
        # (iv) exclusivity: granted >= q and rejected >= q never both
        granted = (cur["votes"] == 1).sum(axis=0)
        rejected = (cur["votes"] == 2).sum(axis=0)
        assert not ((granted >= _q(R)) & (rejected >= _q(R))).any()
        # (v) at most one vote per (group, term)
        same_term = cur["term"] == prev["term"]
        had_vote = prev["vote"] != 0
        assert (cur["vote"][same_term & had_vote] == prev["vote"][same_term & had_vote]).all()
        # leaders: self match == last_index, lead == self
        assert (cur["match"][cur["self_id"][lead] - 1, np.nonzero(lead)[0]] == cur["last_index"][lead]).all()
        assert (cur["lead"][lead] == cur["self_id"][lead]).all()
        prev = cur
    assert o.errors == 0
    assert saw_leader


def test_batched_tick_equals_one_message_at_a_time():
    """(vii): orc_tick over a dense inbox == Step()ping the same messages individually in canonical order
    (senders ascending, then proposals, then a timers-only tick)."""
    G, R, T = 256, 5, 120
    p = _preset(5)
    a = Oracle(G, R, seed=99)
    b = Oracle(G, R, seed=99)
    MsgProp = 2
    for t in range(T):
        ib = a.gen_trace(p, t)
        a.tick(ib)
        for g in range(G):
            b.clear_out(g)
        sid = b.export()["self_id"]
        for r in range(R):
            for g in np.nonzero(ib["type"][r] & 0x0F)[0]:
                if sid[g] == r + 1:
                    continue
                ty = int(ib["type"][r, g])
                b.step(int(g), ty & 0x0F, frm=r + 1, term=int(ib["term"][r, g]), index=int(ib["index"][r, g]),
                       logterm=int(ib["logterm"][r, g]), commit=int(ib["commit"][r, g]), reject=bool(ty & 0x80))
        # NB: per group the order is still sender-ascending; groups are independent, so looping r outside g
        # is the same serialisation per group.
        for g in np.nonzero(ib["prop_count"])[0]:
            b.step(int(g), MsgProp, frm=int(sid[g]), n_entries=int(ib["prop_count"][g]))
        keep = b.export()["out"].copy()
        # timers-only tick must not clear the out word we are comparing: tick() resets out, so compare
        # state columns and OR the out words.
        b.tick(None)
        sa, sb = a.export(), b.export()
        for k in oracle.STATE_COLUMNS:
            np.testing.assert_array_equal(sa[k], sb[k], err_msg=f"tick {t} column {k}")
        np.testing.assert_array_equal(sa["out"], keep | sb["out"])


def test_multithreaded_driver_is_identical():
    G, R, T = 4096, 3, 60
    p = _preset(2)
    a = Oracle(G, R, seed=5)
    b = Oracle(G, R, seed=5)
    for t in range(T):
        ia = a.gen_trace(p, t)
        ib = b.gen_trace(p, t, nthreads=4)
        for k in ia:
            np.testing.assert_array_equal(ia[k], ib[k])
        a.tick(ia, nthreads=1)
        b.tick(ib, nthreads=4)
    sa, sb = a.export(), b.export()
    for k in oracle.STATE_COLUMNS + ("out",):
        np.testing.assert_array_equal(sa[k], sb[k])
    assert (sa["role"] == LEADER).mean() > 0.5


def test_export_import_roundtrip():
    G, R = 700, 5
    p = _preset(5)
    a = Oracle(G, R, seed=3)
    for t in range(80):
        a.tick(a.gen_trace(p, t))
    s = a.export()
    b = Oracle(G, R, seed=3)
    b.import_state(s)
    b.tick_count = a.tick_count
    for t in range(80, 140):
        ia = a.gen_trace(p, t)
        ib = b.gen_trace(p, t)
        for k in ia:
            np.testing.assert_array_equal(ia[k], ib[k])
        a.tick(ia)
        b.tick(ib)
    sa, sb = a.export(), b.export()
    for k in oracle.STATE_COLUMNS:
        np.testing.assert_array_equal(sa[k], sb[k], err_msg=k)
