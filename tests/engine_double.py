"""TEST-ONLY double of the GPU engine for rehearsing host-side code paths on the CPU: the CPU oracle behind the
Engine methods and the few raw C-ABI calls that bench.py and the byte-form GPU tests make.  Posted byte-form frames
are decoded with `mrq_unpack8` — the same inline codec the device kernel runs (include/mrq_packed8.h).

It exists to separate two kinds of failure before code meets hardware: a rehearsal that passes here shows the
host-side code (tests, bench legs) is sound, so a failure on the GPU implicates the device path alone.  It proves
nothing about the device kernels themselves."""
import ctypes as C

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


def _apply_wide(cols, wide):
    for g, frm, ty, term, index, logterm, commit in wide:
        r = frm - 1
        cols["type"][r, g], cols["term"][r, g], cols["index"][r, g] = ty, term, index
        cols["logterm"][r, g], cols["commit"][r, g] = logterm, commit


class FakeL:
    """the C-ABI calls made directly (not through Engine methods), over the double's state"""

    def __init__(self, eng):
        self.e = eng
        self.posts = 0

    def mrq_post_inbox_packed(self, h, slot, ref):
        e, v = self.e, ref._obj
        assert v.word_bits == 8
        G, R = e.G, e.R
        word = np.ctypeslib.as_array((C.c_uint8 * max(1, (R - 1) * G)).from_address(v.word))[: (R - 1) * G].reshape(R - 1, G)
        wide = [(v.wide[i].group, v.wide[i].from_, v.wide[i].type, v.wide[i].term, v.wide[i].index, v.wide[i].logterm,
                 v.wide[i].commit) for i in range(v.n_wide)]
        prop8 = np.ctypeslib.as_array(v.prop_count8, shape=(G,)) if v.prop_count8 else None
        e.post_inbox_packed(word, prop8, wide, slot=slot)
        self.posts += 1
        return 0

    def mrq_tick(self, h, slot):
        self.e.tick(slot)
        return 0

    def mrq_drain_commit_deltas(self, h, dptr):
        e = self.e
        c = e.o.export()["committed"]
        d = c - e.prev
        np.ctypeslib.as_array(dptr, shape=(e.G,))[:] = np.minimum(d, 255).astype(np.uint8)
        e.prev = np.where(d > 255, e.prev, c)
        return 0

    def mrq_drain_tick_deltas(self, h, dptr):
        np.ctypeslib.as_array(dptr, shape=(self.e.G,))[:] = self.e.last_adv
        return 0

    def mrq_drain_wait(self, h):
        return 0

    def mrq_last_error(self, h):
        return b""


class FakeEngine:
    def __init__(self, G, R, seed=0, group_base=0, device=0, inbox_slots=2, self_id=0, election_tick=10, heartbeat_tick=1, **_):
        self.G, self.R = G, R
        self.o = Oracle(G, R, seed=seed, group_base=group_base, self_id=self_id, election_tick=election_tick,
                        heartbeat_tick=heartbeat_tick)
        self.slots = [oracle.empty_inbox(G, R) for _ in range(inbox_slots)]
        self.base_index = np.zeros(G, np.uint64)
        self.base_term = np.zeros(G, np.uint64)
        self.prev = np.zeros(G, np.uint64)
        self.L, self.h = FakeL(self), 1
        self.tick_mode, self.frames = 0, {}
        self.last_adv, self.slot_out = np.zeros(G, np.uint8), {}

    def _self_id(self):
        return self.o.export()["self_id"]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def import_state(self, st):
        self.o.import_state(st)

    def export_state(self, columns=None):
        return self.o.export()

    def gen_trace(self, p, t, slot=0):
        self.frames.pop(slot, None)
        q = oracle.TraceParams()
        for n, _ in F.TraceParams._fields_:
            setattr(q, n, getattr(p, n))
        self.slots[slot] = self.o.gen_trace(q, t)

    def read_inbox(self, slot=0):
        return {k: v.copy() for k, v in self.slots[slot].items()}

    def post_inbox_packed(self, word, prop8=None, wide=(), slot=0, keep=False):
        assert word.dtype == np.uint8 and word.shape == (max(self.R - 1, 0), self.G)
        frame = (np.array(word), None if prop8 is None else np.array(prop8), list(wide), bool(keep))
        if self.tick_mode >= 3:  # the frame waits in its slot; the TICK decodes it, against the base of that moment
            self.frames[slot] = frame
        else:  # the unpack pass runs at post time
            self.frames.pop(slot, None)
            self._decode(frame, slot)

    def _decode(self, frame, slot):
        word, prop8, wide, _ = frame
        cols, self.base_index = unpack8(word, self._self_id(), self.base_index, self.base_term, self.R)
        _apply_wide(cols, wide)
        cols["prop_count"] = np.zeros(self.G, np.uint32) if prop8 is None else np.asarray(prop8).astype(np.uint32)
        self.slots[slot] = cols

    def post_inbox_delta(self, msgs, slot=0, accumulate=False):
        self.frames.pop(slot, None)
        if not accumulate:
            self.slots[slot] = oracle.empty_inbox(self.G, self.R)
        _apply_wide(self.slots[slot], msgs)

    def propose(self, groups, counts, slot=0):
        for g, c in zip(groups, counts):
            self.slots[slot]["prop_count"][g] += c

    def clear_inbox(self, slot=0):
        self.slots[slot] = oracle.empty_inbox(self.G, self.R)

    def tick(self, slot=0):
        if self.tick_mode >= 3 and slot in self.frames:
            frame = self.frames[slot]
            self._decode(frame, slot)
            if not frame[3]:  # MRQ_PACKED_KEEP not set: one tick per post
                del self.frames[slot]
        before = self.o.export()["committed"]
        self.o.tick(self.slots[slot])
        after = self.o.export()
        self.last_adv = np.minimum(after["committed"] - before, 255).astype(np.uint8)  # mode 4's per-tick advance bytes
        self.slot_out[slot] = (after["out"].copy(), self.last_adv.copy())

    def tick_many(self, slots):
        for s in slots:
            self.tick(s)

    def sync_tick_deltas(self):
        return self.last_adv.copy()

    def sync_slot_outputs(self, slot):
        return self.slot_out[slot]

    def set_write_through(self, on):
        pass

    def post_inbox_dense(self, ib, slot=0):
        self.frames.pop(slot, None)
        self.slots[slot] = {k: np.array(v) for k, v in ib.items()}

    def match_update(self, groups, frm, index):
        st = self.o.export()
        for g, f, i in zip(groups, frm, index):
            self.o.step(int(g), 4, frm=int(f), term=int(st["term"][int(g)]), index=int(i))
            self._undo_commit = True

    def quorum_commit(self):
        pass  # (the double's match_update already Stepped the acks one at a time, commit included)

    def tick_idle(self, n=1):
        for _ in range(n):
            self.o.tick(None)

    def sync_out(self):
        return self.o.export()["out"]

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
        if bi is not None:
            self.base_index = np.array(bi, np.uint64)
        if bt is not None:
            self.base_term = np.array(bt, np.uint64)

    def synchronize(self):
        pass

    def set_tick_mode(self, mode):
        self.tick_mode = mode  # every mode computes the same thing; mode 3 differs in WHEN a byte frame is decoded

    def counters(self):
        return {"errors": self.o.errors, "ticks": self.o.tick_count}

    def close(self):
        pass


class FakeBenchEngine(FakeEngine):
    """+ the calls bench.py's main path makes around its timed region (no timing meaning: the double only lets the
    orchestration and the JSON assembly run on the CPU)"""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.launches = 0
        self._t0 = 0.0

    def set_graph_mode(self, mode):
        pass

    def set_l2_policy(self, on):
        pass

    def tick(self, slot=0):
        super().tick(slot)
        self.launches += 2

    def counters(self):
        return {"kernel_launches": self.launches, "ticks": self.o.tick_count, "errors": self.o.errors}

    def timer_start(self):
        import time

        self._t0 = time.perf_counter()

    def timer_stop(self):
        import time

        return 1e3 * (time.perf_counter() - self._t0)
