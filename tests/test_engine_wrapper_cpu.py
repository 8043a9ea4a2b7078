"""The REAL raftsql_b200.engine.Engine wrappers against a stub of libmrq (no GPU, no library call): every argument
the wrapper marshals is read back from the ctypes structures the stub receives.  tests/engine_double.py re-implements
Engine for the host-harness rehearsals, so a bug in the real wrapper's marshalling (round 1: a local shadowing the
`keep` flag of post_inbox_packed) was invisible to the CPU suite; this file closes that hole."""
import ctypes as C
import inspect

import numpy as np
import pytest

from raftsql_b200 import _ffi as F
from raftsql_b200.engine import Engine, MrqError


class StubLib:
    """records (name, args); struct arguments arrive as byref() objects whose ._obj is the structure"""

    def __init__(self, rc=0):
        self.calls, self.rc = [], rc

    def __getattr__(self, name):
        if name not in F.SIGNATURES:
            raise AttributeError(name)
        res = F.SIGNATURES[name][0]

        def fn(*args):
            assert len(args) == len(F.SIGNATURES[name][1]), f"{name}: arity"
            self.calls.append((name, args))
            if name == "mrq_last_error":
                return b"stub error"
            if res is C.c_int:
                return self.rc
            return 0

        return fn


def make(G, R, rc=0):
    e = Engine.__new__(Engine)
    e.L, e.h, e.G, e.R = StubLib(rc), C.c_void_p(0x1234), G, R
    return e


def last(e, name):
    for n, a in reversed(e.L.calls):
        if n == name:
            return a
    raise AssertionError(f"{name} was not called")


@pytest.mark.parametrize("dtype,bits", [(np.uint32, 32), (np.uint16, 16), (np.uint8, 8)])
@pytest.mark.parametrize("keep", [False, True])
def test_post_inbox_packed_marshals_word_bits_and_keep_flag(dtype, bits, keep):
    G, R = 37, 5
    rows = R - 1 if bits == 8 else R
    word = (np.arange(rows * G).reshape(rows, G) % 251).astype(dtype)
    prop8 = np.arange(G, dtype=np.uint8)
    wide = [(3, 2, F.MSG_VOTE, 9, 100, 8, 0), (36, 5, F.MSG_APP, 9, 101, 9, 77)]
    e = make(G, R)
    e.post_inbox_packed(word, prop8, wide, slot=2, keep=keep)
    h, slot, ref = last(e, "mrq_post_inbox_packed")
    v = ref._obj
    assert slot == 2 and v.word == word.ctypes.data and v.word_bits == bits
    assert v.reserved == (1 if keep else 0)
    assert C.addressof(v.prop_count8.contents) == prop8.ctypes.data
    assert v.n_wide == 2 and (v.wide[0].group, v.wide[0].from_, v.wide[0].type, v.wide[0].term) == (3, 2, F.MSG_VOTE, 9)
    assert (v.wide[1].index, v.wide[1].logterm, v.wide[1].commit) == (101, 9, 77)
    assert e.L.calls[-1][0] == "mrq_synchronize"  # numpy temporaries are safe on return


def test_post_inbox_packed_byte_form_with_one_replica_has_no_sender_rows():
    e = make(64, 1)
    word = np.zeros((0, 64), np.uint8)
    e.post_inbox_packed(word, np.zeros(64, np.uint8), keep=True)
    v = last(e, "mrq_post_inbox_packed")[2]._obj
    assert v.word and v.word_bits == 8 and v.reserved == 1 and v.n_wide == 0


def test_post_inbox_packed_rejects_the_wrong_shape():
    e = make(16, 3)
    with pytest.raises(AssertionError):
        e.post_inbox_packed(np.zeros((3, 16), np.uint8))  # byte form has R-1 rows
    with pytest.raises(AssertionError):
        e.post_inbox_packed(np.zeros((2, 16), np.uint16))


def test_error_codes_become_exceptions_with_the_library_message():
    e = make(8, 3, rc=F.MRQ_E_INVAL)
    with pytest.raises(MrqError) as ei:
        e.set_tick_mode(3)
    assert ei.value.code == F.MRQ_E_INVAL and "stub error" in str(ei.value)


def test_every_simple_wrapper_forwards_its_arguments():
    G, R = 10, 3
    e = make(G, R)
    e.set_tick_mode(3)
    assert last(e, "mrq_set_tick_mode")[1] == 3
    e.tick(4)
    assert last(e, "mrq_tick")[1] == 4
    e.tick_many([0, 1, 2])
    a = last(e, "mrq_tick_many")
    assert a[2] == 3 and [a[1][i] for i in range(3)] == [0, 1, 2]
    e.tick_idle(7)
    assert last(e, "mrq_tick_idle")[1] == 7
    e.set_quorum_variant(2)
    assert last(e, "mrq_set_quorum_variant")[1] == 2
    e.propose([1, 5], [2, 3], slot=1)
    a = last(e, "mrq_propose")
    assert a[1] == 1 and a[4] == 2 and (a[2][0], a[2][1], a[3][0], a[3][1]) == (1, 5, 2, 3)
    e.match_update([4], [2], [99])
    a = last(e, "mrq_match_update")
    assert (a[1][0], a[2][0], a[3][0], a[4]) == (4, 2, 99, 1)
    e.post_inbox_delta([(1, 2, F.MSG_APP_RESP, 3, 4, 0, 0)], slot=1, accumulate=True)
    a = last(e, "mrq_post_inbox_delta")
    assert a[1] == 1 and a[3] == 1 and a[4] == 1 and (a[2][0].group, a[2][0].from_, a[2][0].index) == (1, 2, 4)
    bi, bt = np.arange(G, dtype=np.uint64), np.ones(G, np.uint64)
    e.set_packed_base(bi, bt)
    a = last(e, "mrq_set_packed_base")
    assert C.addressof(a[1].contents) == bi.ctypes.data and C.addressof(a[2].contents) == bt.ctypes.data
    ib = dict(type=np.zeros((R, G), np.uint8), term=np.zeros((R, G), np.uint64), index=np.zeros((R, G), np.uint64),
              logterm=np.zeros((R, G), np.uint64), commit=np.zeros((R, G), np.uint64), prop_count=np.zeros(G, np.uint32))
    e.post_inbox_dense(ib, slot=1)
    v = last(e, "mrq_post_inbox_dense")[2]._obj
    assert C.addressof(v.index.contents) == ib["index"].ctypes.data and C.addressof(v.prop_count.contents) == ib["prop_count"].ctypes.data
    st = e.export_state(columns=("term", "match"))
    v = last(e, "mrq_export_state")[1]._obj
    assert C.addressof(v.term.contents) == st["term"].ctypes.data and not v.vote and st["match"].shape == (R, G)
    e.import_state({"committed": np.arange(G, dtype=np.uint64)})
    v = last(e, "mrq_import_state")[1]._obj
    assert v.committed[3] == 3 and not v.term
    c, r, t = e.sync_commits(want_role=True, want_term=True)
    assert c.shape == (G,) and r.dtype == np.uint8 and t.dtype == np.uint64
    assert e.sync_out().dtype == np.uint32 and e.sync_commit_deltas().dtype == np.uint8


def test_mode_4_wrappers_forward_their_arguments():
    e = make(12, 5)
    e.set_tick_mode(4)
    assert last(e, "mrq_set_tick_mode")[1] == 4
    e.set_write_through(0)
    assert last(e, "mrq_set_write_through")[1] == 0
    d = e.sync_tick_deltas()
    a = last(e, "mrq_drain_tick_deltas")
    assert d.dtype == np.uint8 and d.shape == (12,) and C.addressof(a[1].contents) == d.ctypes.data
    names = [n for n, _ in e.L.calls]
    assert names.index("mrq_drain_tick_deltas") < names.index("mrq_drain_wait")  # the copy is waited for before the array is used
    o, dl = e.sync_slot_outputs(3)
    a = last(e, "mrq_sync_slot_outputs")
    assert a[1] == 3 and o.dtype == np.uint32 and dl.dtype == np.uint8
    assert C.addressof(a[2].contents) == o.ctypes.data and C.addressof(a[3].contents) == dl.ctypes.data


def test_the_engine_double_has_the_same_method_signatures_as_the_real_wrapper():
    """the rehearsal double must not drift from the wrapper it stands in for"""
    from engine_double import FakeEngine as EngineDouble

    for name, fn in inspect.getmembers(EngineDouble, inspect.isfunction):
        if name.startswith("_") or not hasattr(Engine, name):
            continue
        real = getattr(Engine, name)
        if isinstance(real, property):
            continue
        ps_d = [p for p in inspect.signature(fn).parameters]
        ps_r = [p for p in inspect.signature(real).parameters]
        # the double may accept fewer options, but every parameter it has must exist, in order, on the real wrapper
        assert len(ps_d) <= len(ps_r), f"{name}: double {ps_d} vs real {ps_r}"
        for i, q in enumerate(ps_d):  # the names callers pass by keyword must sit where the real wrapper has them
            if q in ("slot", "keep", "accumulate", "wide", "prop8", "columns"):
                assert ps_r[i] == q, f"{name}: double {ps_d} vs real {ps_r}"
