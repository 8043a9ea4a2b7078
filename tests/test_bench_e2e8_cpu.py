"""bench.py's byte-form end-to-end leg (`run_e2e8_child`) rehearsed on the CPU: the REAL orchestration code — frame
building with `mrq_pack8` as the trace is generated, the pipelined post / tick / drain loop, the commit-advance
accumulation and the equality verdict — driven against a test double of the engine: the CPU oracle behind the
handful of C-ABI calls the leg makes, decoding posted frames with `mrq_unpack8` (the same inline decode the device
kernel runs).  What this cannot cover is the device kernel itself and real PCIe timing; what it does cover is
every line of the leg's Python and the claim it checks at run time (byte-form ticks commit exactly what the wide
inbox commits)."""
import argparse
import ctypes as C
import json

import numpy as np

import oracle
from oracle import Oracle
from raftsql_b200 import _ffi as F
from raftsql_b200.packed import unpack8


class FakePinned:
    def __init__(self, shape, dtype):
        self.array = np.zeros(shape, dtype)
        self.ptr = self.array.ctypes.data
        self.nbytes = self.array.nbytes

    def free(self):
        pass


class FakeL:
    """the C-ABI calls the leg makes directly, over the double's state"""

    def __init__(self, eng):
        self.e = eng
        self.posts = 0

    def mrq_post_inbox_packed(self, h, slot, ref):
        e, v = self.e, ref._obj
        assert v.word_bits == 8
        G, R = e.G, e.R
        word = np.ctypeslib.as_array((C.c_uint8 * ((R - 1) * G)).from_address(v.word)).reshape(R - 1, G)
        cols, e.base_index = unpack8(word, e.self_id, e.base_index, e.base_term, R)
        for i in range(v.n_wide):
            m = v.wide[i]
            r, g = m.from_ - 1, m.group
            cols["type"][r, g], cols["term"][r, g], cols["index"][r, g] = m.type, m.term, m.index
            cols["logterm"][r, g], cols["commit"][r, g] = m.logterm, m.commit
        cols["prop_count"] = np.ctypeslib.as_array(v.prop_count8, shape=(G,)).astype(np.uint32)
        e.slots[slot] = cols
        self.posts += 1
        return 0

    def mrq_tick(self, h, slot):
        self.e.o.tick(self.e.slots[slot])
        return 0

    def mrq_drain_commit_deltas(self, h, dptr):
        e = self.e
        c = e.o.export()["committed"]
        d = c - e.prev
        np.ctypeslib.as_array(dptr, shape=(e.G,))[:] = np.minimum(d, 255).astype(np.uint8)
        e.prev = np.where(d > 255, e.prev, c)
        return 0

    def mrq_drain_wait(self, h):
        return 0

    def mrq_last_error(self, h):
        return b""


class FakeEngine:
    def __init__(self, G, R, seed=0, group_base=0, device=0, inbox_slots=2):
        self.G, self.R = G, R
        self.o = Oracle(G, R, seed=seed, group_base=group_base)
        self.slots = [oracle.empty_inbox(G, R) for _ in range(inbox_slots)]
        self.base_index = np.zeros(G, np.uint64)
        self.base_term = np.zeros(G, np.uint64)
        self.prev = np.zeros(G, np.uint64)
        self.self_id = self.o.export()["self_id"].copy()
        self.L, self.h = FakeL(self), 1

    def import_state(self, st):
        self.o.import_state(st)
        self.self_id = self.o.export()["self_id"].copy()

    def gen_trace(self, p, t, slot=0):
        q = oracle.TraceParams()
        for n, _ in F.TraceParams._fields_:
            setattr(q, n, getattr(p, n))
        self.slots[slot] = self.o.gen_trace(q, t)

    def read_inbox(self, slot=0):
        return {k: v.copy() for k, v in self.slots[slot].items()}

    def tick(self, slot=0):
        self.o.tick(self.slots[slot])

    def sync_commits(self):
        c = self.o.export()["committed"].copy()
        self.prev = c.copy()  # a full read rebases the delta drain (mrq_sync_commits)
        return c

    @property
    def tick_count(self):
        return self.o.tick_count

    @tick_count.setter
    def tick_count(self, t):
        self.o.tick_count = t

    def set_packed_base(self, bi, bt):
        self.base_index, self.base_term = np.array(bi, np.uint64), np.array(bt, np.uint64)

    def synchronize(self):
        pass

    def close(self):
        pass


def test_byte_form_leg_end_to_end_on_the_double(monkeypatch, capsys):
    import bench
    import raftsql_b200
    import raftsql_b200.packed as packed

    made = []

    def make(*a, **kw):
        made.append(FakeEngine(*a, **kw))
        return made[-1]

    monkeypatch.setattr(bench, "G_TOTAL", 6000)
    monkeypatch.setattr(raftsql_b200, "Engine", make)
    monkeypatch.setattr(packed, "PinnedArray", FakePinned)
    bench.run_e2e8_child(argparse.Namespace(steps=7))
    res = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert res["equals_wide_form"] is True
    assert res["steps"] == 7 and res["escapes"] == 0 and res["value"] > 0
    assert res["h2d_bytes_per_step"] == 6000 * bench.R  # (R-1) sender bytes + 1 proposal byte per group
    assert res["d2h_bytes_per_step"] == 6000
    assert made[0].L.posts == 3 + 7 + 7  # the three runs of the leg really went through the byte-form post
    # and the verdict is not vacuous: the trace commits entries on every one of those ticks
    st = made[0].o.export()
    assert (st["committed"] > bench.steady_state(6000, bench.R, 0, bench.SEED)["committed"]).mean() > 0.9
