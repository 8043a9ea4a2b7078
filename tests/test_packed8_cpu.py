"""The byte form of the packed inbox (include/mrq_packed8.h): codec, row mapping, sliding window and the host
frame builder `mrq_pack8`, all CPU code in libmrq.so — no device needed.  `mrq_unpack8` runs the very inline
decode the device kernel runs, so  unpack8(pack8(x)) + scatter(wide) == x  here is the exactness claim of the
form; the device side of it (same decode inside unpack8_inbox_kernel) is in tests/test_zz_packed8_gpu.py."""
import numpy as np
import pytest

import oracle
from oracle import Oracle, TraceParams
from raftsql_b200 import _ffi as F
from raftsql_b200.packed import Pack8, unpack8

ESC = 0xFE


def rebuild(decoded, wide):
    """decoded columns + the escape list -> what the tick will see (wide messages override their slot)."""
    out = {k: v.copy() for k, v in decoded.items()}
    for g, frm, ty, term, index, logterm, commit in wide:
        r = frm - 1
        out["type"][r, g], out["term"][r, g], out["index"][r, g] = ty, term, index
        out["logterm"][r, g], out["commit"][r, g] = logterm, commit
    return out


def assert_same_messages(ib, got, self_id):
    """Equal as far as Step() can tell: type everywhere (own-slot cells are dropped by both), and for every
    present message the fields its type carries (include/mrq.h, 'Field meaning per type')."""
    R, G = ib["type"].shape
    own = (np.arange(R)[:, None] + 1) == self_id[None, :]
    want_type = np.where(own, 0, ib["type"])
    assert np.array_equal(got["type"], want_type)
    kind = want_type & 0x0F
    present = kind != 0
    assert np.array_equal(got["term"][present], ib["term"][present])
    for types, col in (((F.MSG_APP_RESP, F.MSG_VOTE, F.MSG_APP), "index"), ((F.MSG_VOTE, F.MSG_APP), "logterm"),
                       ((F.MSG_HEARTBEAT, F.MSG_APP), "commit")):
        m = np.isin(kind, types)
        assert np.array_equal(got[col][m], ib[col][m]), col


def test_every_byte_decodes_to_what_the_header_says():
    base_i, base_t = np.array([1000], np.uint64), np.array([7], np.uint64)
    for b in range(256):
        for self_id in (1, 2):
            word = np.array([[b]], np.uint8)
            out, nb = unpack8(word, np.array([self_id], np.uint8), base_i, base_t, 2)
            r = 1 if self_id == 1 else 0  # the only remote sender
            kind, p = b & 3, b >> 2
            ty = int(out["type"][r, 0])
            assert out["type"][1 - r, 0] == 0
            if kind == 1:
                assert ty == F.MSG_APP_RESP and out["index"][r, 0] == 1000 + p and out["term"][r, 0] == 7
                assert nb[0] == (1000 + p - 16 if p > 16 else 1000)
            elif kind == 3:
                assert ty == F.MSG_HEARTBEAT and out["commit"][r, 0] == 1000 + p and out["term"][r, 0] == 7
                assert nb[0] == 1000
            elif kind == 2 and p in (0, 1, 2):
                assert ty == {0: F.MSG_HEARTBEAT_RESP, 1: F.MSG_VOTE_RESP, 2: F.MSG_VOTE_RESP | F.MSG_REJECT}[p]
                assert out["term"][r, 0] == 7 and nb[0] == 1000
            else:  # none, escape, never-produced payloads
                assert ty == 0 and nb[0] == 1000


@pytest.mark.parametrize("R", [1, 2, 3, 5, 8])
def test_round_trip_on_message_soup(R):
    """every type, terms around the base, indices inside / at the edges of / outside the window, any self id"""
    rng = np.random.default_rng(R)
    G = 3000
    self_id = rng.integers(1, R + 1, size=G).astype(np.uint8)
    base_i = rng.integers(0, 1 << 40, size=G).astype(np.uint64)
    base_i[:50] = rng.integers(0, 20, size=50)  # near zero: index < base must escape, not wrap
    base_t = rng.integers(1, 9, size=G).astype(np.uint64)
    pk = Pack8(self_id, base_i, base_t, R)
    types = np.array([0, 0, 3, 4, 4, 4, 4, 5, 6, 6 | 0x80, 8, 8, 9, 4 | 0x80, 3 | 0x80, 9 | 0x80, 8 | 0x80], np.uint8)
    for _ in range(12):
        ib = oracle.empty_inbox(G, R)
        ib["type"][:] = rng.choice(types, size=(R, G))
        ib["term"][:] = (pk.base_term[None, :].astype(np.int64) + rng.choice(np.array([0] * 9 + [-1, 1]), size=(R, G))).astype(np.uint64)
        off = rng.choice(np.array([-2, -1, 0, 1, 15, 16, 17, 30, 62, 63, 64, 65, 1 << 33]), size=(R, G))
        ib["index"][:] = np.maximum(pk.base_index[None, :].astype(np.int64) + off, 0).astype(np.uint64)
        ib["commit"][:] = np.maximum(pk.base_index[None, :].astype(np.int64) + np.roll(off, 1, axis=1), 0).astype(np.uint64)
        ib["logterm"][:] = rng.integers(0, 9, size=(R, G)).astype(np.uint64)
        ib["prop_count"][:] = rng.integers(0, 256, size=G).astype(np.uint32)
        before = pk.base_index.copy()
        word, p8, wide = pk.frame(ib)
        assert word.shape == (max(R - 1, 0), G)
        assert np.array_equal(p8, ib["prop_count"].astype(np.uint8))
        decoded, slid = unpack8(word, self_id, before, base_t, R)
        assert np.array_equal(slid, pk.base_index), "host and device windows must slide identically"
        assert_same_messages(ib, rebuild(decoded, wide), self_id)
        # escapes are exactly the bytes marked so, one wide message each
        assert int((word == ESC).sum()) == len(wide)
        # and nothing that fits was escaped: in-window acks / heartbeats of the base term never ride wide
        for g, frm, ty, term, index, logterm, commit in wide[:200]:
            if ty == F.MSG_APP_RESP and term == base_t[g]:
                assert not (0 <= index - int(before[g]) <= 63)
            if ty == F.MSG_HEARTBEAT and term == base_t[g]:
                assert not (0 <= commit - int(before[g]) <= 63)


def test_window_follows_a_steady_state_leader_without_escapes():
    """BASELINE configs[2] shape: every follower acks every tick, the log grows ~1.5 entries per tick.  After the
    initial placement the window must keep up by itself: no escapes, no host rebase, for hundreds of ticks."""
    G, R, T = 2048, 5, 400
    rng = np.random.default_rng(3)
    o = Oracle(G, R, seed=11)
    st = o.export()
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
    st["term_start"][:] = st["committed"] - np.uint64(5)
    st["randomized_timeout"][:] = 10
    o.import_state(st)
    p = TraceParams()
    p.seed, p.p_ack_256, p.p_grant_256, p.max_prop, p.lag_kind = 0x5EED0003, 256, 230, 3, 1
    pk = Pack8(st["self_id"], st["last_index"] - np.uint64(40), st["term"], R)
    dev_base = pk.base_index.copy()
    escapes = 0
    for t in range(T):
        ib = o.gen_trace(p, t)
        word, p8, wide = pk.frame(ib)
        escapes += len(wide)
        decoded, dev_base = unpack8(word, st["self_id"], dev_base, st["term"], R)
        assert np.array_equal(dev_base, pk.base_index)
        assert_same_messages(ib, rebuild(decoded, wide), st["self_id"])
        o.tick(ib)
    assert escapes == 0
    assert (pk.base_index > st["last_index"]).all(), "the window must have travelled with the log"
    assert word.nbytes + p8.nbytes == G * R  # (R-1) sender bytes + 1 proposal byte per group


@pytest.mark.parametrize("G,R", [(1, 2), (7, 3), (1000, 5), (33333, 8), (65, 1)])
def test_frame_builder_writes_only_inside_its_output_buffers(G, R):
    """canaries around word_out / prop8_out / the decode's output columns: mrq_pack8 and mrq_unpack8 must not write a
    byte outside [R-1][G] / [G] / [R][G] (the library is C: numpy would not notice)"""
    rng = np.random.default_rng(G + R)
    PAD = 4096

    def guarded(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        raw = np.full(n + 2 * PAD, 0xA5, np.uint8)
        return raw, raw[PAD:PAD + n].view(dtype).reshape(shape)

    self_id = rng.integers(1, R + 1, size=G).astype(np.uint8)
    base_i = rng.integers(100, 1 << 40, size=G).astype(np.uint64)
    base_t = rng.integers(1, 9, size=G).astype(np.uint64)
    ib = oracle.empty_inbox(G, R)
    ib["type"][:] = rng.choice(np.array([0, 4, 4, 4, 8, 9, 6, 5, 3], np.uint8), size=(R, G))
    ib["term"][:] = base_t[None, :]
    ib["index"][:] = base_i[None, :] + rng.integers(0, 80, size=(R, G)).astype(np.uint64)
    ib["commit"][:] = base_i[None, :] + rng.integers(0, 80, size=(R, G)).astype(np.uint64)
    ib["prop_count"][:] = rng.integers(0, 200, size=G)
    raw_w, word = guarded((max(R - 1, 0), G), np.uint8)
    raw_p, p8 = guarded((G,), np.uint8)
    pk = Pack8(self_id, base_i, base_t, R)
    before = pk.base_index.copy()
    for _ in range(3):
        _, _, wide = pk.frame(ib, word_out=word, prop8_out=p8)
    for raw in (raw_w, raw_p):
        assert (raw[:PAD] == 0xA5).all() and (raw[-PAD:] == 0xA5).all(), "mrq_pack8 wrote outside its output buffer"
    # the decode side, through the raw ABI with guarded columns
    L = F.load()
    cols, raws = {}, []
    for k, dt in (("type", np.uint8), ("term", np.uint64), ("index", np.uint64), ("logterm", np.uint64), ("commit", np.uint64)):
        raw, cols[k] = guarded((R, G), dt)
        raws.append(raw)
    raw_b, base = guarded((G,), np.uint64)
    base[:] = before
    import ctypes as C

    view = F.InboxOut(*(cols[k].ctypes.data_as(t) for k, t in (("type", F.u8p), ("term", F.u64p), ("index", F.u64p),
                                                              ("logterm", F.u64p), ("commit", F.u64p))), None)
    word0, _, _ = Pack8(self_id, before, base_t, R).frame(ib)
    rc = L.mrq_unpack8(word0.ctypes.data_as(F.u8p) if word0.size else None, self_id.ctypes.data_as(F.u8p), G, R,
                       base.ctypes.data_as(F.u64p), base_t.ctypes.data_as(F.u64p), C.byref(view))
    assert rc == F.MRQ_OK
    for raw in raws + [raw_b]:
        assert (raw[:PAD] == 0xA5).all() and (raw[-PAD:] == 0xA5).all(), "mrq_unpack8 wrote outside its output buffer"


def test_frame_is_the_same_on_one_and_on_many_host_threads(monkeypatch):
    """mrq_pack8 splits large frames into contiguous group ranges on host threads: bytes, escape order (group
    order) and the slid bases must not depend on how many."""
    rng = np.random.default_rng(8)
    G, R = 200_000, 5
    self_id = (np.arange(G) % R + 1).astype(np.uint8)
    li = rng.integers(1 << 20, 1 << 40, size=G).astype(np.uint64)
    bt = rng.integers(1, 9, size=G).astype(np.uint64)
    ib = oracle.empty_inbox(G, R)
    ib["type"][:] = F.MSG_APP_RESP
    ib["term"][:] = bt[None, :]
    ib["index"][:] = li[None, :] - rng.geometric(0.2, size=(R, G)).astype(np.uint64)
    ib["index"][2, ::97] += np.uint64(1 << 20)  # escapes scattered over every chunk
    ib["term"][4, ::1013] += np.uint64(1)
    ib["prop_count"][:] = rng.integers(0, 4, size=G)
    got = {}
    for nt in ("1", "6"):
        monkeypatch.setenv("MRQ_HOST_THREADS", nt)
        pk = Pack8(self_id, li - np.uint64(40), bt, R)
        w1, p1, wide1 = pk.frame(ib)
        w2, p2, wide2 = pk.frame(ib)  # a second frame from the slid base
        got[nt] = (w1.tobytes(), p1.tobytes(), wide1, w2.tobytes(), wide2, pk.base_index.tobytes())
    assert got["1"] == got["6"]
    assert len(got["1"][2]) > 1500 and [m[0] for m in got["1"][2]] == sorted(m[0] for m in got["1"][2])


def test_too_many_proposals_and_small_escape_buffers_are_refused_without_side_effects():
    L = F.load()
    G, R = 4, 3
    ib = oracle.empty_inbox(G, R)
    ib["type"][1, :] = F.MSG_VOTE  # always escapes
    ib["term"][1, :] = 5
    self_id = np.ones(G, np.uint8)
    pk = Pack8(self_id, np.zeros(G, np.uint64), np.full(G, 5, np.uint64), R)
    word, p8, wide = pk.frame(ib)  # the wrapper grows the escape buffer on demand
    assert len(wide) == G and all(m[2] == F.MSG_VOTE and m[1] == 2 for m in wide)
    ib["prop_count"][2] = 256
    with pytest.raises(ValueError, match="255"):
        pk.frame(ib)
    assert L.mrq_pack8(None, None, 0, 3, None, None, None, None, None, 0, None) == F.MRQ_E_INVAL
