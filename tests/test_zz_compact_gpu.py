"""Tick mode 4 on hardware: the tick on COMPACT state (32-bit offsets from a per-group base, four groups per thread)
reading the byte inbox itself — per-tick launches and mrq_tick_many's single launch over a whole slot sequence.
The per-group arithmetic (compact_step / materialise_group / compact_group / slow_group_ticks_c) is verified on the host
against the CPU checker (tests/cpp/tick_host_test.cpp, "compact ..."); here the vector loads / stores, the slow list,
the per-slot outputs, the staging-buffer lifetime and every conversion between the two representations meet the GPU:
state, out words and commit advances must equal the oracle's, bit for bit."""
import numpy as np
import pytest

import oracle
from oracle import Oracle
from raftsql_b200 import Engine, preset_trace
from raftsql_b200 import _ffi as F
from raftsql_b200.packed import Pack8
from util import assert_state_equal

pytestmark = [pytest.mark.gpu]


@pytest.fixture(params=["4 groups per thread", "1 group per thread"], autouse=True)
def both_kernel_shapes(request, monkeypatch):
    """the library picks the quad kernel for big shards and the one-group-per-thread kernel for small ones; the tests
    here are small, so each runs once under either shape (MRQ_T4_GPT is read at every launch)"""
    monkeypatch.setenv("MRQ_T4_GPT", "4" if request.param.startswith("4") else "1")


def _orc_params(p):
    q = oracle.TraceParams()
    for n, _ in F.TraceParams._fields_:
        setattr(q, n, getattr(p, n))
    return q


def _warm(G, R, seed, ticks, cfg_no, slots=3):
    p = preset_trace(cfg_no)
    eng, orc = Engine(G, R, seed=seed, inbox_slots=slots), Oracle(G, R, seed=seed)
    for t in range(ticks):
        eng.gen_trace(p, t)
        ib = eng.read_inbox()
        eng.tick()
        orc.tick(ib)
    return eng, orc, p


def _rebase(eng, orc, self_id, R, back=30):
    c = orc.export()
    pk = Pack8(self_id, np.where(c["last_index"] > back, c["last_index"] - np.uint64(back), 0).astype(np.uint64), c["term"], R)
    eng.set_packed_base(pk.base_index, pk.base_term)
    return pk


@pytest.mark.parametrize("G,R,cfg", [(4000, 7, 5), (3001, 5, 3), (777, 2, 5), (64, 1, 2), (2500, 8, 5), (3000, 3, 2), (5, 5, 5), (513, 4, 5)])
def test_mode_4_per_tick_launches_equal_the_oracle(G, R, cfg):
    eng, orc, p = _warm(G, R, 51 + R, 60, cfg)
    eng.set_tick_mode(4)
    self_id = orc.export()["self_id"].copy()
    pk = _rebase(eng, orc, self_id, R)
    for t in range(60, 170):
        if t % 25 == 0:
            pk = _rebase(eng, orc, self_id, R)
        ib = orc.gen_trace(_orc_params(p), t)
        word, prop8, wide = pk.frame(ib)
        before = orc.export()["committed"].copy()
        eng.post_inbox_packed(word, prop8, wide, slot=t % 3)
        eng.tick(t % 3)
        orc.tick(ib)
        want = orc.export()
        np.testing.assert_array_equal(eng.sync_out(), want["out"], err_msg=f"out word, tick {t}")
        np.testing.assert_array_equal(eng.sync_tick_deltas(), np.minimum(want["committed"] - before, 255).astype(np.uint8),
                                      err_msg=f"commit advance, tick {t}")
        if t % 7 == 0 or t > 160:  # (export converts compact -> wide: do it on some ticks only, so runs of compact ticks happen)
            assert_state_equal(eng.export_state(), want, f"mode 4 tick {t}")
    assert_state_equal(eng.export_state(), orc.export(), "mode 4 final")
    c = eng.counters()
    assert c["errors"] == 0 and orc.errors == 0
    eng.close()


@pytest.mark.parametrize("G,R,cfg,K,wt", [(4000, 5, 5, 6, 1), (3001, 5, 3, 12, 0), (2049, 7, 5, 5, 0), (1000, 3, 2, 8, 1), (600, 2, 5, 3, 0)])
def test_mode_4_tick_many_runs_a_whole_slot_sequence_in_one_launch(G, R, cfg, K, wt):
    """K frames posted to K slots, then ONE mrq_tick_many: every tick's out words and commit advances (per-slot buffers)
    and the final state must equal the oracle's; write-through on and off."""
    eng, orc, p = _warm(G, R, 61 + R, 60, cfg, slots=K)
    eng.set_tick_mode(4)
    eng.set_write_through(wt)
    self_id = orc.export()["self_id"].copy()
    pk = _rebase(eng, orc, self_id, R)
    t = 60
    for batch in range(10):
        if batch % 3 == 2:
            pk = _rebase(eng, orc, self_id, R)
        want_out, want_adv = [], []
        for k in range(K):
            ib = orc.gen_trace(_orc_params(p), t)
            word, prop8, wide = pk.frame(ib)
            eng.post_inbox_packed(word, prop8, wide, slot=k)
            before = orc.export()["committed"].copy()
            orc.tick(ib)
            after = orc.export()
            want_out.append(after["out"].copy())
            want_adv.append(np.minimum(after["committed"] - before, 255).astype(np.uint8))
            t += 1
        eng.tick_many(list(range(K)))
        for k in range(K):
            o, d = eng.sync_slot_outputs(k)
            np.testing.assert_array_equal(o, want_out[k], err_msg=f"out words of tick {k} of batch {batch}")
            np.testing.assert_array_equal(d, want_adv[k], err_msg=f"commit advances of tick {k} of batch {batch}")
        np.testing.assert_array_equal(eng.sync_out(), want_out[-1])
        assert_state_equal(eng.export_state(), orc.export(), f"after batch {batch}")
    assert eng.tick_count == t
    c = eng.counters()
    assert c["errors"] == 0 and orc.errors == 0
    eng.close()


def test_mode_4_steady_state_window_slides_for_hundreds_of_ticks_in_batches():
    """bench shape: steady-state leaders, one set_packed_base, then 320 frames in batches of 16 with kept frames"""
    G, R, K = 8192, 5, 16
    rng = np.random.default_rng(5)
    eng, orc = Engine(G, R, seed=77, inbox_slots=K), Oracle(G, R, seed=77)
    st = orc.export()
    g = np.arange(G)
    st["self_id"][:] = (g % R + 1).astype(np.uint8)
    st["role"][:] = 2
    st["lead"][:] = st["self_id"]
    st["term"][:] = rng.integers(1, 9, size=G).astype(np.uint64)
    st["vote"][:] = st["self_id"]
    st["last_index"][:] = rng.integers(1 << 20, 1 << 40, size=G).astype(np.uint64)
    st["last_term"][:] = st["term"]
    st["match"][:] = st["last_index"][None, :] - rng.geometric(0.2, size=(R, G)).astype(np.uint64)
    st["match"][st["self_id"] - 1, g] = st["last_index"]
    st["committed"][:] = st["last_index"] - np.uint64(40)
    st["term_start"][:] = np.where(rng.random(G) < 0.9, st["committed"] - np.uint64(5), st["last_index"] - np.uint64(1))
    st["randomized_timeout"][:] = 10
    orc.import_state(st)
    eng.import_state(st)
    eng.set_tick_mode(4)
    p = preset_trace(3)
    base = (st["last_index"] - np.uint64(40)).astype(np.uint64)
    eng.set_packed_base(base, st["term"])
    pk = Pack8(st["self_id"], base, st["term"], R)
    t = 0
    n_commits = 0
    for batch in range(20):
        for k in range(K):
            ib = orc.gen_trace(_orc_params(p), t)
            word, prop8, wide = pk.frame(ib)
            assert not wide
            eng.post_inbox_packed(word, prop8, wide, slot=k, keep=True)
            orc.tick(ib)
            n_commits += int(((orc.export()["out"] & F.OUT_COMMIT_ADVANCED) != 0).sum())
            t += 1
        eng.tick_many(list(range(K)))
        if batch % 5 == 4:
            assert_state_equal(eng.export_state(), orc.export(), f"batch {batch}")
    assert_state_equal(eng.export_state(), orc.export(), "final")
    c = eng.counters()
    assert c["errors"] == 0 and c["commits_advanced"] == n_commits and n_commits > G * t // 2
    eng.close()


def test_mode_4_interoperates_with_every_entry_point_that_touches_wide_state():
    """export / partial import / a wide dense post / an idle tick / the standalone quorum kernel / match_update / a mode
    switch in the middle of a mode-4 run: each converts between the representations, none may change a result."""
    G, R = 3000, 5
    eng, orc, p = _warm(G, R, 9, 50, 5)
    eng.set_tick_mode(4)
    self_id = orc.export()["self_id"].copy()
    pk = _rebase(eng, orc, self_id, R)
    po = _orc_params(p)

    def byte_tick(t):
        ib = orc.gen_trace(po, t)
        word, prop8, wide = pk.frame(ib)
        eng.post_inbox_packed(word, prop8, wide, slot=0)
        eng.tick(0)
        orc.tick(ib)

    t = 50
    for _ in range(5):
        byte_tick(t)
        t += 1
    assert_state_equal(eng.export_state(), orc.export(), "a. after byte ticks")
    # b. a wide (dense) post in mode 4 ticks through the general kernels on the wide columns
    ib = orc.gen_trace(po, t)
    eng.post_inbox_dense(ib, slot=1)
    eng.tick(1)
    orc.tick(ib)
    t += 1
    assert_state_equal(eng.export_state(), orc.export(), "b. wide post in mode 4")
    pk = _rebase(eng, orc, self_id, R)  # the wide tick moved the log without the frame builder: re-base it
    for _ in range(3):
        byte_tick(t)
        t += 1
    # c. an idle tick (timers only)
    eng.tick_idle(1)
    orc.tick(None)
    t += 1
    assert_state_equal(eng.export_state(), orc.export(), "c. idle tick")
    for _ in range(3):
        byte_tick(t)
        t += 1
    # d. a partial import (committed only) must leave every other column exact
    cur = orc.export()
    lead = cur["role"] == 2
    newc = cur["committed"].copy()
    orc.import_state({"committed": newc})
    eng.import_state({"committed": newc})
    for _ in range(3):
        byte_tick(t)
        t += 1
    assert_state_equal(eng.export_state(), orc.export(), "d. partial import")
    # e. sparse maybeUpdate + the standalone quorum kernel on wide state, then back to byte ticks
    es = eng.export_state()
    gs = np.nonzero(lead)[0].astype(np.uint64)
    fr = ((es["self_id"][lead] % R) + 1).astype(np.uint8)
    idx = es["last_index"][lead]
    eng.match_update(gs, fr, idx)
    for g_, f_, i_ in zip(gs, fr, idx):
        orc.step(int(g_), 4, frm=int(f_), term=int(es["term"][g_]), index=int(i_))
    eng.quorum_commit()
    np.testing.assert_array_equal(eng.sync_commits(), orc.export()["committed"])
    pk = _rebase(eng, orc, self_id, R)
    for _ in range(4):
        byte_tick(t)
        t += 1
    assert_state_equal(eng.export_state(), orc.export(), "e. K3 in between")
    # f. leave mode 4 mid-run (mode 3 reads the same frames on wide state), come back
    eng.set_tick_mode(3)
    for _ in range(3):
        byte_tick(t)
        t += 1
    assert_state_equal(eng.export_state(), orc.export(), "f. mode 3 in between")
    eng.set_tick_mode(4)
    for _ in range(3):
        byte_tick(t)
        t += 1
    assert_state_equal(eng.export_state(), orc.export(), "f. back in mode 4")
    assert eng.counters()["errors"] == 0 and orc.errors == 0
    eng.close()


def test_mode_4_at_the_benchmark_shape_equals_mode_0():
    """1,048,576 x 5 steady state (bench.py's workload): 12 ticks in mode 4 (6 per-tick, 6 in one launch) against the same
    ticks from the wide inbox in mode 0 — every state column, and the commit indices reconstructed from the byte drain."""
    import bench

    G, R = 1 << 20, 5
    st0 = bench.steady_state(G, R, 0, bench.SEED)
    p = preset_trace(3)
    T = 12
    ref = Engine(G, R, seed=bench.SEED, inbox_slots=2)
    ref.import_state(st0)
    eng = Engine(G, R, seed=bench.SEED, inbox_slots=T)
    eng.import_state(st0)
    eng.set_tick_mode(4)
    base0 = (st0["last_index"] - np.uint64(40)).astype(np.uint64)
    eng.set_packed_base(base0, st0["term"])
    pk = Pack8(st0["self_id"], base0, st0["term"], R)
    acc = st0["committed"].copy()
    for t in range(T):
        ref.gen_trace(p, t, slot=0)
        ib = ref.read_inbox(0)
        ref.tick(0)
        word, prop8, wide = pk.frame(ib)
        assert not wide
        eng.post_inbox_packed(word, prop8, wide, slot=t, keep=True)
        if t < 6:
            eng.tick(t)
            d = eng.sync_tick_deltas()
            assert d.max() < 255
            acc += d
    np.testing.assert_array_equal(acc, eng.sync_commits())
    eng.tick_many(list(range(6, T)))
    for k in range(6, T):
        _, d = eng.sync_slot_outputs(k)
        acc += d
    want = ref.export_state()
    got = eng.export_state()
    np.testing.assert_array_equal(acc, want["committed"])
    for k in ("term", "vote", "committed", "last_index", "last_term", "term_start", "match", "role", "lead", "election_elapsed",
              "heartbeat_elapsed"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    np.testing.assert_array_equal(eng.sync_out(), ref.sync_out())
    c = eng.counters()
    assert c["errors"] == 0 and c["commits_advanced"] == ref.counters()["commits_advanced"]
    ref.close()
    eng.close()


def test_mode_4_many_distinct_slot_sequences_stay_exact():
    """mrq_tick_many caches one descriptor table per slot sequence and drops the cache when it holds 64: a host that never
    repeats a sequence (here 80 different ones over 5 slots) must get the same results as the oracle throughout"""
    import itertools

    G, R = 700, 3
    eng, orc, p = _warm(G, R, 17, 50, 2, slots=5)
    eng.set_tick_mode(4)
    self_id = orc.export()["self_id"].copy()
    pk = _rebase(eng, orc, self_id, R)
    po = _orc_params(p)
    t = 50
    seqs = [list(s) for n in (2, 3, 4, 5) for s in itertools.permutations(range(5), n)][:80]
    assert len({tuple(s) for s in seqs}) == 80
    for k, seq in enumerate(seqs):
        if k % 10 == 9:
            pk = _rebase(eng, orc, self_id, R)
        for slot in seq:
            ib = orc.gen_trace(po, t)
            word, prop8, wide = pk.frame(ib)
            eng.post_inbox_packed(word, prop8, wide, slot=slot)
            orc.tick(ib)
            t += 1
        eng.tick_many(seq)
        if k % 8 == 7:
            assert_state_equal(eng.export_state(), orc.export(), f"sequence {k}")
    assert_state_equal(eng.export_state(), orc.export(), "final")
    assert eng.tick_count == t and eng.counters()["errors"] == 0
    eng.close()
