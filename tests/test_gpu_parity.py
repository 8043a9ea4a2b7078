"""GPU parity: the sm_100a engine (through the C-ABI) against the CPU oracle, bit for bit.

Every test drives libmrq.so exactly as a host would — post an inbox, tick, drain — and compares raw
uint64 / uint8 arrays with the oracle's (memcmp semantics, no tolerance: this is integer work).
"""
import numpy as np
import pytest

import oracle
from oracle import Oracle
from raftsql_b200 import Engine, empty_inbox, preset_trace
from raftsql_b200 import _ffi as F
from util import (LEADER, U64MAX, assert_inbox_equal, assert_state_equal, leader_state, numpy_quorum_index)

pytestmark = pytest.mark.gpu


def _orc_params(p):
    q = oracle.TraceParams()
    for n, _ in F.TraceParams._fields_:
        setattr(q, n, getattr(p, n))
    return q


def run_trace_parity(G, R, cfg_no, T, *, seed, check_every=1, nthreads=1, self_id=0, group_base=0):
    p = preset_trace(cfg_no)
    eng = Engine(G, R, seed=seed, self_id=self_id, group_base=group_base)
    orc = Oracle(G, R, seed=seed, self_id=self_id, group_base=group_base)
    assert_state_equal(eng.export_state(), orc.export(), "initial")
    for t in range(T):
        eng.gen_trace(p, t, slot=0)
        ib = eng.read_inbox(0)
        if t % check_every == 0:  # device generator == host generator on the same state
            assert_inbox_equal(ib, orc.gen_trace(_orc_params(p), t, nthreads=nthreads), f"tick {t}")
        eng.tick(0)
        orc.tick(ib, nthreads=nthreads)
        if t % check_every == 0 or t == T - 1:
            assert_state_equal(eng.export_state(), orc.export(), f"tick {t}")
            np.testing.assert_array_equal(eng.sync_out(), orc.export()["out"], err_msg=f"out word, tick {t}")
    c = eng.counters()
    assert c["errors"] == 0 and orc.errors == 0
    s = orc.export()
    eng.close()
    return c, s


def test_config2_4096x3_every_tick():
    """BASELINE configs[1]: 4,096 groups x 3 replicas, election + replication trace, every tick compared."""
    c, s = run_trace_parity(4096, 3, 2, 1024, seed=0x5EED0002)
    assert (s["role"] == LEADER).mean() > 0.95  # the trace really elects leaders and replicates
    assert c["elections_won"] >= 4096 * 0.95 and c["commits_advanced"] > 4096 * 100
    assert (s["committed"] > 100).mean() > 0.9


def test_config5_lag_and_churn_small_every_tick():
    """BASELINE configs[4] shape (7 replicas, 20% lagging followers, leader churn), 8,192 groups."""
    c, s = run_trace_parity(8192, 7, 5, 700, seed=0x5EED0005)
    assert c["step_downs"] > 1000 and c["campaigns"] > 8192 and c["votes_granted"] > 100


def test_config5_full_size_262144x7():
    """BASELINE configs[4] at full size; state compared every 16th tick and at the end."""
    c, s = run_trace_parity(262144, 7, 5, 160, seed=0x5EED0005, check_every=16, nthreads=oracle.hw_threads())
    assert c["step_downs"] > 10000


def test_config5_spec_length_262144x7_4096_ticks_compared_every_tick():
    """BASELINE configs[4] exactly as SURVEY §8d states it: 262,144 groups x 7 replicas, 20 % lagging followers, leader
    churn, 4,096 ticks, `term / vote / role / committed` bit-equal to the oracle after EVERY tick.  Engine and oracle each
    generate the tick's trace from their own state (same generator: include/mrq_trace.h), so a divergence of either
    shows at once; the full state, the out word and the generated inbox are compared every 256th tick as well."""
    G, R, T = 262144, 7, 4096
    nt = oracle.hw_threads()
    p = preset_trace(5)
    po = _orc_params(p)
    eng = Engine(G, R, seed=0x5EED0005)
    orc = Oracle(G, R, seed=0x5EED0005)
    cols = ("term", "vote", "role", "committed")
    for t in range(T):
        eng.gen_trace(p, t, slot=0)
        if t % 256 == 0:
            assert_inbox_equal(eng.read_inbox(0), orc.gen_trace(po, t, nthreads=nt), f"tick {t}")
        eng.tick(0)
        orc.tick(orc.gen_trace(po, t, nthreads=nt), nthreads=nt)
        es, os_ = eng.export_state(columns=cols), orc.export()
        for k in cols:
            if not np.array_equal(es[k], os_[k]):
                bad = np.flatnonzero(es[k] != os_[k])
                raise AssertionError(f"tick {t}: column {k} differs on {len(bad)} groups, first {bad[0]}: "
                                     f"engine {es[k][bad[0]]} oracle {os_[k][bad[0]]}")
        if t % 256 == 255 or t == T - 1:
            assert_state_equal(eng.export_state(), os_, f"tick {t}")
            np.testing.assert_array_equal(eng.sync_out(), os_["out"], err_msg=f"out word, tick {t}")
    c = eng.counters()
    s = orc.export()
    assert c["errors"] == 0 and orc.errors == 0
    assert c["step_downs"] > 200000 and c["elections_won"] > 200000 and (s["role"] == LEADER).mean() > 0.5
    eng.close()


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 6, 7, 8])
def test_all_replica_counts(R):
    run_trace_parity(1000, R, 5, 200, seed=77 + R)


@pytest.mark.parametrize("G", [1, 2, 31, 63, 64, 65, 255, 257])
def test_ragged_group_counts(G):
    run_trace_parity(G, 5, 5, 120, seed=G)


def test_fixed_self_id_and_group_base():
    run_trace_parity(2048, 5, 5, 150, seed=9, self_id=3)
    run_trace_parity(2048, 3, 2, 150, seed=9, group_base=1 << 33)


def _follower_heavy_params():
    """Leaders get deposed by higher-term heartbeats and then keep hearing from that leader: most groups end
    up as followers receiving same-term heartbeats (the follower half of the fast kernel)."""
    p = preset_trace(5)
    p.churn_65536, p.p_heartbeat_256, p.p_grant_256, p.p_reject_256 = 600, 235, 150, 60
    return p


@pytest.mark.parametrize("R", [3, 5, 7])
def test_follower_heartbeat_trace(R):
    G, T = 6000, 260
    p = _follower_heavy_params()
    eng, orc = Engine(G, R, seed=41), Oracle(G, R, seed=41)
    for t in range(T):
        eng.gen_trace(p, t)
        ib = eng.read_inbox()
        eng.tick()
        orc.tick(ib)
        if t % 4 == 0 or t == T - 1:
            assert_state_equal(eng.export_state(), orc.export(), f"tick {t}")
            np.testing.assert_array_equal(eng.sync_out(), orc.export()["out"])
    s = orc.export()
    assert ((s["role"] == 0) & (s["lead"] != 0)).mean() > 0.3  # plenty of followers with a live leader
    assert eng.counters()["errors"] == 0 and orc.errors == 0


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("cfg_no,R", [(2, 3), (5, 5), (5, 7), (3, 5)])
def test_split_launch_equals_single_general_kernel(cfg_no, R, mode):
    """mrq_set_tick_mode: fast + slow kernels (0) / the fused single launch (2) vs one general kernel over
    every group (1)."""
    G, T = 20000, 150
    p = preset_trace(cfg_no) if cfg_no != 3 else _follower_heavy_params()
    a, b = Engine(G, R, seed=7), Engine(G, R, seed=7)
    a.set_tick_mode(mode)
    b.set_tick_mode(1)
    if cfg_no == 3:
        st = leader_state(G, R, np.random.default_rng(1))
        a.import_state(st)
        b.import_state(st)
    for t in range(T):
        a.gen_trace(p, t)
        ib = a.read_inbox()
        b.post_inbox_dense(ib)
        a.tick()
        b.tick()
        if t % 10 == 0 or t == T - 1:
            sa, sb = a.export_state(), b.export_state()
            lead = sa["role"] == LEADER
            for k in sa:
                if k == "match":
                    np.testing.assert_array_equal(sa[k][:, lead], sb[k][:, lead], err_msg=f"tick {t} match")
                else:
                    np.testing.assert_array_equal(sa[k], sb[k], err_msg=f"tick {t} {k}")
            np.testing.assert_array_equal(a.sync_out(), b.sync_out())
    ca, cb = a.counters(), b.counters()
    for k in ("campaigns", "elections_won", "step_downs", "commits_advanced", "votes_granted", "errors"):
        assert ca[k] == cb[k], k


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tick_many_graph_replay_equals_single_ticks(mode):
    """mrq_tick_many (CUDA-graph replay of a slot sequence) == the same ticks launched one by one, including a
    second replay of the cached graph and a rewind (set_tick_count normalises the graph parities)."""
    G, R, S = 30000, 5, 7
    p = preset_trace(5)
    a, b = Engine(G, R, seed=3, inbox_slots=S), Engine(G, R, seed=3, inbox_slots=S)
    a.set_tick_mode(mode)
    b.set_tick_mode(mode)
    for t in range(40):  # get past the first elections
        a.gen_trace(p, t)
        b.post_inbox_dense(a.read_inbox())
        a.tick()
        b.tick()
    snap, tick0 = a.export_state(), a.tick_count
    for rep in range(3):
        for s in range(S):  # S inboxes generated from a's evolving state, mirrored into b's slots
            a.gen_trace(p, 100 + rep * S + s, slot=s)
            b.post_inbox_dense(a.read_inbox(s), slot=s)
            a.tick(s)
        b.tick_many(list(range(S)))
        sa, sb = a.export_state(), b.export_state()
        lead = sa["role"] == LEADER
        for k in sa:
            if k == "match":
                np.testing.assert_array_equal(sa[k][:, lead], sb[k][:, lead], err_msg=f"rep {rep} match")
            else:
                np.testing.assert_array_equal(sa[k], sb[k], err_msg=f"rep {rep} {k}")
        assert a.tick_count == b.tick_count
    # rewind both and replay the last slot sequence: the cached graph must still be valid
    for e in (a, b):
        e.import_state(snap)
        e.tick_count = tick0
    for s in range(S):
        a.tick(s)
    b.tick_many(list(range(S)))
    sa, sb = a.export_state(), b.export_state()
    for k in ("term", "committed", "last_index", "role", "votes", "randomized_timeout", "election_elapsed"):
        np.testing.assert_array_equal(sa[k], sb[k], err_msg=f"after rewind {k}")
    assert a.counters()["commits_advanced"] == b.counters()["commits_advanced"]


def test_zero_groups():
    with Engine(0, 3) as eng:
        eng.tick_idle(3)
        eng.quorum_commit()
        assert eng.tick_count == 3
        assert eng.sync_commits().shape == (0,)


def test_idle_ticks_only_timers():
    G, R = 3000, 3
    eng, orc = Engine(G, R, seed=4), Oracle(G, R, seed=4)
    for t in range(45):
        eng.tick_idle(1)
        orc.tick(None)
        assert_state_equal(eng.export_state(), orc.export(), f"idle tick {t}")
    s = orc.export()
    assert (s["role"] == 1).all() and (s["term"] >= 2).all()  # everyone timed out and campaigned, twice


def test_config3_steady_state_1Mx5_import_and_tick():
    """BASELINE configs[2]: 1,048,576 x 5 steady state imported, then the append/ack trace."""
    G, R = 1 << 20, 5
    rng = np.random.default_rng(3)
    st = leader_state(G, R, rng)
    eng, orc = Engine(G, R, seed=0x5EED0003), Oracle(G, R, seed=0x5EED0003)
    eng.import_state(st)
    orc.import_state(st)
    assert_state_equal(eng.export_state(), orc.export(), "import")
    p = preset_trace(3)
    nt = oracle.hw_threads()
    for t in range(6):
        eng.gen_trace(p, t)
        ib = eng.read_inbox()
        eng.tick()
        orc.tick(ib, nthreads=nt)
        assert_state_equal(eng.export_state(), orc.export(), f"tick {t}")
    assert eng.counters()["commits_advanced"] > G


# ---- K3: the standalone quorum kernel ------------------------------------------------------------------

@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 6, 7, 8])
def test_quorum_kernel_vs_oracle_and_numpy(R, variant):
    G = 70001  # odd, not a multiple of the TMA tile: exercises the tiled body and the LDG tail
    rng = np.random.default_rng(1000 + R)
    st = leader_state(G, R, rng, gate_open_frac=0.9)
    # adversarial cells: ties, zeros, huge values, non-leaders (gate closed for ever)
    st["match"][:, :200] = rng.integers(0, 3, size=(R, 200), dtype=np.uint64)
    st["committed"][:200] = 0
    st["term_start"][:200] = rng.integers(0, 3, size=200, dtype=np.uint64)
    st["match"][:, 200:300] = U64MAX - rng.integers(0, 2, size=(R, 100), dtype=np.uint64)
    st["last_index"][200:300] = U64MAX
    st["committed"][200:300] = 5
    st["term_start"][200:300] = 6
    st["role"][300:400] = 0
    st["term_start"][300:400] = U64MAX
    eng, orc = Engine(G, R), Oracle(G, R)
    eng.import_state(st)
    orc.import_state(st)
    eng.set_quorum_variant(variant)
    eng.quorum_commit()
    orc.quorum_commit()
    got = eng.sync_commits()
    np.testing.assert_array_equal(got, orc.export()["committed"])
    mci = numpy_quorum_index(st["match"])
    want = np.where((mci > st["committed"]) & (mci >= st["term_start"]), mci, st["committed"])
    np.testing.assert_array_equal(got, want)
    assert eng.counters()["commits_advanced"] == int((want != st["committed"]).sum())
    # idempotent: a second pass moves nothing
    eng.quorum_commit()
    np.testing.assert_array_equal(eng.sync_commits(), want)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_quorum_kernel_full_size_properties(variant):
    """1,048,576 x 5 (BASELINE configs[2]): q-th largest by an independent numpy sort, monotone, gated,
    idempotent, and linear in a uniform index shift."""
    G, R = 1 << 20, 5
    rng = np.random.default_rng(5)
    st = leader_state(G, R, rng)
    with Engine(G, R) as eng:
        eng.import_state(st)
        eng.set_quorum_variant(variant)
        eng.quorum_commit()
        got = eng.sync_commits()
        mci = numpy_quorum_index(st["match"])
        want = np.where((mci > st["committed"]) & (mci >= st["term_start"]), mci, st["committed"])
        np.testing.assert_array_equal(got, want)
        assert (got >= st["committed"]).all()
        moved = got > st["committed"]
        assert moved.mean() > 0.9 and (got[moved] >= st["term_start"][moved]).all()
        eng.quorum_commit()
        np.testing.assert_array_equal(eng.sync_commits(), want)
        # shift every index by a constant: the result shifts by the same constant
        k = np.uint64(12345)
        st2 = dict(st)
        for col in ("match", "committed", "term_start", "last_index"):
            st2[col] = st[col] + k
        eng.import_state(st2)
        eng.quorum_commit()
        np.testing.assert_array_equal(eng.sync_commits(), want + k)


def test_quorum_ext_on_caller_device_buffers():
    import torch

    G, R, stride = 50000, 5, 50048
    rng = np.random.default_rng(8)
    st = leader_state(G, R, rng)
    m = np.zeros((R, stride), np.uint64)
    m[:, :G] = st["match"]
    dm = torch.from_numpy(m.view(np.int64)).cuda()
    dc = torch.from_numpy(st["committed"].view(np.int64)).cuda()
    dg = torch.from_numpy(st["term_start"].view(np.int64)).cuda()
    with Engine(16, R) as eng:
        for variant in (0, 1, 2):
            c = dc.clone()
            torch.cuda.synchronize()
            eng.quorum_commit_ext(dm.data_ptr(), c.data_ptr(), dg.data_ptr(), G, stride, variant)
            eng.synchronize()
            mci = numpy_quorum_index(st["match"])
            want = np.where((mci > st["committed"]) & (mci >= st["term_start"]), mci, st["committed"])
            np.testing.assert_array_equal(c.cpu().numpy().view(np.uint64), want)


# ---- the other inbox forms, proposals, sparse acks, drains ---------------------------------------------

def _warm(G, R, seed, ticks, cfg_no=5):
    p = preset_trace(cfg_no)
    eng, orc = Engine(G, R, seed=seed, inbox_slots=3), Oracle(G, R, seed=seed)
    for t in range(ticks):
        eng.gen_trace(p, t)
        ib = eng.read_inbox()
        eng.tick()
        orc.tick(ib)
    return eng, orc, p


def test_dense_host_inbox_equals_device_inbox():
    G, R = 5000, 5
    eng, orc, p = _warm(G, R, 21, 60)
    for t in range(60, 120):
        ib = orc.gen_trace(_orc_params(p), t)
        eng.post_inbox_dense(ib, slot=1)
        eng.tick(1)
        orc.tick(ib)
    assert_state_equal(eng.export_state(), orc.export(), "dense host inbox")


def test_sparse_delta_inbox_equals_dense():
    G, R = 3000, 7
    eng, orc, p = _warm(G, R, 22, 50)
    for t in range(50, 90):
        ib = orc.gen_trace(_orc_params(p), t)
        rs, gs = np.nonzero(ib["type"])
        msgs = [(int(g), int(r) + 1, int(ib["type"][r, g]), int(ib["term"][r, g]), int(ib["index"][r, g]),
                 int(ib["logterm"][r, g]), int(ib["commit"][r, g])) for r, g in zip(rs, gs)]
        eng.post_inbox_delta(msgs, slot=2)
        pg = np.nonzero(ib["prop_count"])[0]
        eng.propose(pg, ib["prop_count"][pg], slot=2)
        eng.tick(2)
        orc.tick(ib)
        assert_state_equal(eng.export_state(), orc.export(), f"sparse tick {t}")


@pytest.mark.parametrize("bits", [32, 16])
def test_packed_inbox_equals_dense(bits):
    """The 4- and 2-byte-per-slot host forms decode to exactly the wide inbox (escapes included)."""
    from raftsql_b200.packed import pack_inbox, pack_inbox16

    packer = pack_inbox if bits == 32 else pack_inbox16
    G, R = 4000, 7
    eng, orc, p = _warm(G, R, 26, 60)
    n_escaped = n_packed = 0
    for t in range(60, 130):
        cur = orc.export()
        ib = orc.gen_trace(_orc_params(p), t)
        # bases near the group's state; lagging followers' stale acks fall below the base and must escape
        base_index = np.where(cur["last_index"] > 50, cur["last_index"] - np.uint64(50), 0).astype(np.uint64)
        base_term = np.where(cur["term"] > 0, cur["term"] - np.uint64(t % 2), 0).astype(np.uint64)
        word, prop8, wide = packer(ib, base_index, base_term)
        assert word.dtype == (np.uint32 if bits == 32 else np.uint16)
        n_escaped += len(wide)
        n_packed += int(((word & (15 if bits == 32 else 7)) != 0).sum()) - len(wide)
        eng.set_packed_base(base_index, base_term)
        eng.post_inbox_packed(word, prop8, wide, slot=1)
        got = eng.read_inbox(1)
        kind = ib["type"] & 0x0F
        np.testing.assert_array_equal(got["type"], ib["type"])
        np.testing.assert_array_equal(got["term"][kind != 0], ib["term"][kind != 0])
        # only the columns Step() reads for a type are materialised by the decode
        uses = {"index": (F.MSG_APP_RESP, F.MSG_VOTE, F.MSG_APP), "logterm": (F.MSG_VOTE, F.MSG_APP),
                "commit": (F.MSG_HEARTBEAT, F.MSG_APP)}
        for k, types in uses.items():
            sel = np.isin(kind, types)
            np.testing.assert_array_equal(got[k][sel], ib[k][sel], err_msg=k)
        np.testing.assert_array_equal(got["prop_count"], ib["prop_count"])
        eng.tick(1)
        orc.tick(ib)
        assert_state_equal(eng.export_state(), orc.export(), f"packed tick {t}")
    assert n_escaped > 100 and n_packed > 10 * n_escaped


def test_match_update_then_quorum_equals_step_by_step():
    """a14 as a sparse pass + K3  ==  Step(MsgAppResp) one at a time on the oracle."""
    G, R = 4000, 5
    rng = np.random.default_rng(31)
    st = leader_state(G, R, rng)
    eng, orc = Engine(G, R), Oracle(G, R)
    eng.import_state(st)
    orc.import_state(st)
    n = 6000
    gs = rng.integers(0, G, size=n, dtype=np.uint64)
    fr = rng.integers(1, R + 1, size=n).astype(np.uint8)
    ok = fr != st["self_id"][gs.astype(np.int64)]
    gs, fr = gs[ok], fr[ok]
    idx = st["last_index"][gs.astype(np.int64)] - rng.integers(0, 3, size=len(gs), dtype=np.uint64)
    eng.match_update(gs, fr, idx)
    eng.quorum_commit()
    for g, f, i in zip(gs, fr, idx):
        orc.step(int(g), F.MSG_APP_RESP, frm=int(f), term=int(st["term"][g]), index=int(i))
    orc.quorum_commit()  # K3 is maybeCommit() on every leader, also where no ack raised a match this round
    o = orc.export()
    np.testing.assert_array_equal(eng.sync_commits(), o["committed"])
    np.testing.assert_array_equal(eng.export_state(("match",))["match"], o["match"])


@pytest.mark.parametrize("R", [3, 5, 7])
def test_out_of_range_acks_follow_upstream(R):
    """Acks beyond the leader's lastIndex (a protocol violation): upstream still records the match, and
    term() of such an index is 0 so it never commits — but an earlier in-range quorum index in the same
    tick does.  The engine defers maybeCommit() to once per tick EXCEPT in exactly this case."""
    G = 6000
    rng = np.random.default_rng(900 + R)
    st = leader_state(G, R, rng)
    eng, orc = Engine(G, R, seed=1), Oracle(G, R, seed=1)
    eng.import_state(st)
    orc.import_state(st)
    for t in range(8):
        cur = orc.export()
        ib = empty_inbox(G, R)
        kind = rng.integers(0, 5, size=(R, G))  # 0 none, 1-2 in-range ack, 3 out-of-range ack, 4 stale ack
        li = cur["last_index"][None, :]
        idx = np.where(kind == 3, li + rng.integers(1, 1000, size=(R, G), dtype=np.uint64),
                       np.where(kind == 4, li - np.uint64(30), li - rng.integers(0, 4, size=(R, G), dtype=np.uint64)))
        ib["type"][:] = np.where(kind > 0, F.MSG_APP_RESP, 0).astype(np.uint8)
        ib["term"][:] = np.where(kind > 0, cur["term"][None, :], 0)
        ib["index"][:] = np.where(kind > 0, idx, 0)
        ib["prop_count"][:] = rng.integers(0, 3, size=G, dtype=np.uint32)
        if t == 5:  # a higher-term heartbeat right after the acks: the deferred evaluation must land first
            ib["type"][R - 1, ::7] = F.MSG_HEARTBEAT
            ib["term"][R - 1, ::7] = cur["term"][::7] + np.uint64(1)
            ib["index"][R - 1, ::7] = 0
        eng.post_inbox_dense(ib)
        eng.tick()
        orc.tick(ib)
        assert_state_equal(eng.export_state(), orc.export(), f"tick {t}")
        np.testing.assert_array_equal(eng.sync_out(), orc.export()["out"])
    # the strict marker survives an export/import round trip (it is recomputed from match vs lastIndex)
    eng2 = Engine(G, R, seed=1)
    eng2.import_state(eng.export_state())
    eng2.tick_count = eng.tick_count
    cur = orc.export()
    ib = empty_inbox(G, R)
    ib["type"][0, :] = F.MSG_APP_RESP
    ib["term"][0, :] = cur["term"]
    ib["index"][0, :] = cur["last_index"]
    for e in (eng, eng2):
        e.post_inbox_dense(ib)
        e.tick()
    orc.tick(ib)
    assert_state_equal(eng.export_state(), orc.export(), "after strict")
    assert_state_equal(eng2.export_state(), orc.export(), "after strict, re-imported")


def test_commit_delta_drain_reconstructs_commits():
    G, R = 6000, 3
    eng, orc, p = _warm(G, R, 23, 40, cfg_no=2)
    base = np.zeros(G, np.uint64)
    for t in range(40, 100):
        eng.gen_trace(p, t)
        eng.tick()
        d = eng.sync_commit_deltas()
        full = eng.sync_commits()
        sat = d == 255
        base[~sat] += d[~sat].astype(np.uint64)
        np.testing.assert_array_equal(base[~sat], full[~sat])
        base[sat] = full[sat]  # 255 = "read in full"; the full read rebases the drain
    assert (base > 0).mean() > 0.5


def test_export_import_roundtrip_and_next():
    G, R = 2500, 5
    eng, orc, p = _warm(G, R, 24, 80)
    s = eng.export_state()
    eng2 = Engine(G, R, seed=24)
    eng2.import_state(s)
    eng2.tick_count = eng.tick_count
    for t in range(80, 110):
        for e in (eng, eng2):
            e.gen_trace(p, t)
            e.tick()
    a, b = eng.export_state(), eng2.export_state()
    lead = a["role"] == LEADER
    for k in a:
        if k == "match":
            np.testing.assert_array_equal(a[k][:, lead], b[k][:, lead])
        else:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    nx = eng.export_next()
    want = np.maximum(a["match"] + np.uint64(1), a["term_start"][None, :])
    np.testing.assert_array_equal(nx[:, lead], want[:, lead])


def test_counters_match_oracle_out_words():
    G, R = 4096, 5
    p = preset_trace(5)
    eng, orc = Engine(G, R, seed=25), Oracle(G, R, seed=25)
    won = stepped = camp = 0
    for t in range(150):
        eng.gen_trace(p, t)
        ib = eng.read_inbox()
        eng.tick()
        orc.tick(ib)
        out = orc.export()["out"]
        won += int(((out & F.OUT_BECAME_LEADER) != 0).sum())
        stepped += int(((out & F.OUT_STEPPED_DOWN) != 0).sum())
        camp += int(((out & F.OUT_CAMPAIGN) != 0).sum())
    c = eng.counters()
    assert c["elections_won"] == won and c["step_downs"] == stepped
    assert c["campaigns"] >= camp  # R=5: every campaign emits MsgVote, so these are equal
    assert c["campaigns"] == camp
    assert c["ticks"] == 150 and c["kernel_launches"] >= 300
