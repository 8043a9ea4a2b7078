"""GPU side of the byte-form inbox (include/mrq_packed8.h, word_bits = 8): unpack8_inbox_kernel must produce exactly
what the host decode `mrq_unpack8` produces (same inline codec; tests/test_packed8_cpu.py proves that one exact),
slide the device's window in step with the frame builder's, and the ticks that consume it must stay bit-equal to
the oracle.  These tests run in-process like every other GPU test: a failure here turns the suite red.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import Oracle
from raftsql_b200 import Engine, preset_trace
from raftsql_b200 import _ffi as F
from raftsql_b200.packed import Pack8, unpack8
from util import assert_state_equal

pytestmark = [pytest.mark.gpu]


def _orc_params(p):
    q = oracle.TraceParams()
    for n, _ in F.TraceParams._fields_:
        setattr(q, n, getattr(p, n))
    return q


def _warm(G, R, seed, ticks, cfg_no):
    p = preset_trace(cfg_no)
    eng, orc = Engine(G, R, seed=seed, inbox_slots=3), Oracle(G, R, seed=seed)
    for t in range(ticks):
        eng.gen_trace(p, t)
        ib = eng.read_inbox()
        eng.tick()
        orc.tick(ib)
    return eng, orc, p


@pytest.mark.parametrize("G,R,cfg", [(4000, 7, 5), (3001, 5, 3), (777, 2, 5), (64, 1, 2), (2500, 8, 5)])
def test_byte_form_decodes_like_the_host_and_ticks_like_the_oracle(G, R, cfg):
    eng, orc, p = _warm(G, R, 31 + R, 60, cfg)
    cur = orc.export()
    self_id = cur["self_id"].copy()
    base_index = np.where(cur["last_index"] > 30, cur["last_index"] - np.uint64(30), 0).astype(np.uint64)
    base_term = cur["term"].copy()
    eng.set_packed_base(base_index, base_term)
    pk = Pack8(self_id, base_index, base_term, R)
    dev_base = base_index.copy()
    n_escaped = n_bytes = 0
    for t in range(60, 140):
        if t % 25 == 0:  # terms moved on under churn: the host re-bases now and then, as a real one would
            cur = orc.export()
            pk = Pack8(self_id, np.where(cur["last_index"] > 30, cur["last_index"] - np.uint64(30), 0).astype(np.uint64), cur["term"], R)
            eng.set_packed_base(pk.base_index, pk.base_term)
            dev_base = pk.base_index.copy()
        ib = orc.gen_trace(_orc_params(p), t)
        word, prop8, wide = pk.frame(ib)
        n_escaped += len(wide)
        n_bytes += int((word != 0).sum()) - len(wide)
        want, dev_base = unpack8(word, self_id, dev_base, pk.base_term, R)
        np.testing.assert_array_equal(dev_base, pk.base_index)
        eng.post_inbox_packed(word, prop8, wide, slot=1)
        got = eng.read_inbox(1)
        # (1) the device decode == the host decode (+ the wide overrides), for the columns Step() reads
        for g, frm, ty, term, index, logterm, commit in wide:
            r = frm - 1
            want["type"][r, g], want["term"][r, g], want["index"][r, g] = ty, term, index
            want["logterm"][r, g], want["commit"][r, g] = logterm, commit
        kind = want["type"] & 0x0F
        np.testing.assert_array_equal(got["type"], want["type"])
        np.testing.assert_array_equal(got["term"][kind != 0], want["term"][kind != 0])
        for k, types in {"index": (F.MSG_APP_RESP, F.MSG_VOTE, F.MSG_APP), "logterm": (F.MSG_VOTE, F.MSG_APP),
                         "commit": (F.MSG_HEARTBEAT, F.MSG_APP)}.items():
            sel = np.isin(kind, types)
            np.testing.assert_array_equal(got[k][sel], want[k][sel], err_msg=k)
        np.testing.assert_array_equal(got["prop_count"], ib["prop_count"])
        # (2) and it is the inbox the trace meant (own-slot cells never exist in the trace)
        np.testing.assert_array_equal(got["type"], ib["type"])
        eng.tick(1)
        orc.tick(ib)
        assert_state_equal(eng.export_state(), orc.export(), f"byte-form tick {t}")
    if R > 1:
        assert n_bytes > 10 * max(1, n_escaped)
    eng.close()


@pytest.mark.parametrize("G,R,cfg", [(4000, 7, 5), (3001, 5, 3), (777, 2, 5), (2500, 8, 5), (3000, 3, 2)])
def test_tick_mode_3_consumes_the_bytes_itself(G, R, cfg):
    """tick mode 3: no unpack pass — tick_fast8_kernel reads the frame in the staging buffer, tick_slow8_kernel
    materialises only the groups the fast kernel declines (escaped senders keep what the wide list scattered).  The
    per-group arithmetic of both is verified on the host (tests/cpp/tick_host_test.cpp, 'byte form direct'); here
    the launch wrappers, the slow list and the staging-buffer lifetime meet hardware: state must equal the oracle's
    after every tick, and the device's window must end where the frame builder's does."""
    eng, orc, p = _warm(G, R, 41 + R, 60, cfg)
    eng.set_tick_mode(3)
    cur = orc.export()
    self_id = cur["self_id"].copy()

    def rebase():
        c = orc.export()
        pk = Pack8(self_id, np.where(c["last_index"] > 30, c["last_index"] - np.uint64(30), 0).astype(np.uint64), c["term"], R)
        eng.set_packed_base(pk.base_index, pk.base_term)
        return pk

    pk = rebase()
    for t in range(60, 160):
        if t % 25 == 0:
            pk = rebase()
        ib = orc.gen_trace(_orc_params(p), t)
        word, prop8, wide = pk.frame(ib)
        eng.post_inbox_packed(word, prop8, wide, slot=t % 3)
        eng.tick(t % 3)
        orc.tick(ib)
        assert_state_equal(eng.export_state(), orc.export(), f"mode 3 tick {t}")
        np.testing.assert_array_equal(eng.sync_out(), orc.export()["out"], err_msg=f"out word, tick {t}")
    c = eng.counters()
    assert c["errors"] == 0 and orc.errors == 0
    eng.close()


def test_device_window_slides_by_itself_for_hundreds_of_ticks():
    """steady-state leaders (bench shape): one set_packed_base, then 300 frames with no host re-base, no escapes"""
    G, R = 8192, 5
    rng = np.random.default_rng(5)
    eng, orc = Engine(G, R, seed=77, inbox_slots=2), Oracle(G, R, seed=77)
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
    st["term_start"][:] = st["committed"] - np.uint64(5)
    st["randomized_timeout"][:] = 10
    orc.import_state(st)
    eng.import_state(st)
    p = preset_trace(3)
    base = (st["last_index"] - np.uint64(40)).astype(np.uint64)
    eng.set_packed_base(base, st["term"])
    pk = Pack8(st["self_id"], base, st["term"], R)
    for t in range(300):
        ib = orc.gen_trace(_orc_params(p), t)
        word, prop8, wide = pk.frame(ib)
        assert not wide
        eng.post_inbox_packed(word, prop8, wide, slot=t % 2)
        eng.tick(t % 2)
        orc.tick(ib)
        if t % 50 == 49:
            assert_state_equal(eng.export_state(), orc.export(), f"tick {t}")
    assert_state_equal(eng.export_state(), orc.export(), "final")
    eng.close()
