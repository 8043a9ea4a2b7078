"""Host-side encoder for the packed inbox (include/mrq.h `mrq_inbox_packed`).

The host→device link (PCIe) is what bounds the end-to-end tick rate, so the host ships one 32-bit word per
(sender, group) slot instead of the 33-byte wide record; the device decodes it exactly against two per-group
base columns (`mrq_set_packed_base`).  Messages that do not fit (far-away indices, terms more than 2 above the
base, MsgApp) ride in the wide escape list — nothing is approximated.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F

PAYLOAD_BITS = 25
PAYLOAD_MAX = (1 << PAYLOAD_BITS) - 1


def pack_inbox(ib: dict, base_index: np.ndarray, base_term: np.ndarray):
    """Encode a wide dense inbox (dict of [R][G] arrays as in `empty_inbox`) into
    (word[R][G] uint32, prop8[G] uint8, wide list of (group, from, type, term, index, logterm, commit))."""
    ty = ib["type"]
    Rr, G = ty.shape
    kind = (ty & F.MSG_TYPE_MASK).astype(np.uint32)
    rej = ((ty & F.MSG_REJECT) != 0).astype(np.uint32)
    bt = base_term[None, :]
    bi = base_index[None, :]
    term, index, logterm, commit = ib["term"], ib["index"], ib["logterm"], ib["commit"]
    present = kind != 0
    tc = term - bt  # wraps for term < base: then it is > 2 and escapes
    ok = present & (tc <= np.uint64(2))
    pay = np.zeros((Rr, G), np.uint64)
    is_ack = kind == F.MSG_APP_RESP
    is_hb = kind == F.MSG_HEARTBEAT
    is_vote = kind == F.MSG_VOTE
    d_idx = index - bi
    d_cm = commit - bi
    d_lt = logterm - bt
    ok &= ~(kind == F.MSG_APP)
    ok &= ~(is_ack & (d_idx > np.uint64(PAYLOAD_MAX)))
    ok &= ~(is_hb & (d_cm > np.uint64(PAYLOAD_MAX)))
    ok &= ~(is_vote & ((d_lt > np.uint64(3)) | (d_idx > np.uint64(PAYLOAD_MAX >> 2))))
    pay = np.where(is_ack, d_idx, pay)
    pay = np.where(is_hb, d_cm, pay)
    pay = np.where(is_vote, (d_idx << np.uint64(2)) | (d_lt & np.uint64(3)), pay)
    word = np.where(ok, kind | (rej << 4) | (tc.astype(np.uint32) << 5) | (pay.astype(np.uint32) << 7), 0)
    esc = present & ~ok
    word = np.where(esc, kind | (np.uint32(3) << 5), word).astype(np.uint32)  # code 3: see the wide list
    wide = [(int(g), int(r) + 1, int(ty[r, g]), int(term[r, g]), int(index[r, g]), int(logterm[r, g]),
             int(commit[r, g])) for r, g in zip(*np.nonzero(esc))]
    prop = ib.get("prop_count")
    prop8 = None
    if prop is not None:
        if prop.max(initial=0) > 255:
            raise ValueError("packed inbox carries at most 255 proposals per group per tick")
        prop8 = prop.astype(np.uint8)
    return np.ascontiguousarray(word), prop8, wide


PAYLOAD16_MAX = (1 << 11) - 1


def pack_inbox16(ib: dict, base_index: np.ndarray, base_term: np.ndarray):
    """The 2-byte-per-slot form (include/mrq.h, word_bits = 16): acks, vote responses, heartbeats and
    heartbeat responses within 2047 of the base; MsgVote, MsgApp and anything out of range escape."""
    ty = ib["type"]
    Rr, G = ty.shape
    kind = (ty & F.MSG_TYPE_MASK).astype(np.uint32)
    rej = (ty & F.MSG_REJECT) != 0
    bt, bi = base_term[None, :], base_index[None, :]
    term, index, commit = ib["term"], ib["index"], ib["commit"]
    present = kind != 0
    tc = term - bt
    is_ack, is_vr = kind == F.MSG_APP_RESP, kind == F.MSG_VOTE_RESP
    is_hb, is_hr = kind == F.MSG_HEARTBEAT, kind == F.MSG_HEARTBEAT_RESP
    d_idx, d_cm = index - bi, commit - bi
    ok = present & (tc <= np.uint64(2)) & (is_ack | is_vr | is_hb | is_hr)
    ok &= ~(is_ack & (d_idx > np.uint64(PAYLOAD16_MAX)))
    ok &= ~(is_hb & (d_cm > np.uint64(PAYLOAD16_MAX)))
    code = np.zeros((Rr, G), np.uint32)
    code = np.where(is_ack, np.where(rej, 2, 1), code)
    code = np.where(is_vr, np.where(rej, 4, 3), code)
    code = np.where(is_hb, 5, code)
    code = np.where(is_hr, 6, code)
    pay = np.where(is_ack, d_idx, np.where(is_hb, d_cm, 0)).astype(np.uint32)
    word = np.where(ok, code | (tc.astype(np.uint32) << 3) | (pay << 5), 0)
    esc = present & ~ok
    word = np.where(esc, 7, word).astype(np.uint16)
    wide = [(int(g), int(r) + 1, int(ty[r, g]), int(term[r, g]), int(index[r, g]), int(ib["logterm"][r, g]),
             int(commit[r, g])) for r, g in zip(*np.nonzero(esc))]
    prop = ib.get("prop_count")
    prop8 = None
    if prop is not None:
        if prop.max(initial=0) > 255:
            raise ValueError("packed inbox carries at most 255 proposals per group per tick")
        prop8 = prop.astype(np.uint8)
    return np.ascontiguousarray(word), prop8, wide


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ct)


class Pack8:
    """Frame builder for the byte form (include/mrq_packed8.h; `mrq_pack8` in libmrq.so — host CPU code).

    Keeps the host's copy of the sliding window base: construct it with the base columns handed to
    `mrq_set_packed_base`, then call `frame()` once per inbox, in posting order."""

    def __init__(self, self_id: np.ndarray, base_index: np.ndarray, base_term: np.ndarray, n_replicas: int):
        self.L = F.load()
        self.R = int(n_replicas)
        self.G = int(self_id.shape[0])
        self.self_id = np.ascontiguousarray(self_id, np.uint8)
        self.base_index = np.ascontiguousarray(base_index, np.uint64).copy()
        self.base_term = np.ascontiguousarray(base_term, np.uint64).copy()
        self._cap, self._wide = 0, (F.Msg * 1)()  # escape list, grown on demand

    def frame(self, ib: dict, word_out: np.ndarray | None = None, prop8_out: np.ndarray | None = None):
        """-> (word[R-1][G] uint8, prop8[G] uint8 or None, wide list of mrq_msg tuples).  Slides the window."""
        G, Rr = self.G, self.R
        cols = {k: np.ascontiguousarray(ib[k], dt) for k, dt in (("type", np.uint8), ("term", np.uint64), ("index", np.uint64),
                                                                  ("logterm", np.uint64), ("commit", np.uint64))}
        assert cols["type"].shape == (Rr, G)
        prop = ib.get("prop_count")
        prop = None if prop is None else np.ascontiguousarray(prop, np.uint32)
        view = F.Inbox(_p(cols["type"], F.u8p), _p(cols["term"], F.u64p), _p(cols["index"], F.u64p), _p(cols["logterm"], F.u64p),
                       _p(cols["commit"], F.u64p), _p(prop, F.u32p) if prop is not None else None)
        word = word_out if word_out is not None else np.zeros((max(Rr - 1, 0), G), np.uint8)
        assert word.dtype == np.uint8 and word.shape == (max(Rr - 1, 0), G) and word.flags.c_contiguous
        p8 = None
        if prop is not None:
            p8 = prop8_out if prop8_out is not None else np.zeros(G, np.uint8)
        n_wide = C.c_size_t(0)
        while True:
            rc = self.L.mrq_pack8(C.byref(view), _p(self.self_id, F.u8p), G, Rr, _p(self.base_index, F.u64p),
                                  _p(self.base_term, F.u64p), _p(word, F.u8p) if word.size else None,
                                  _p(p8, F.u8p) if p8 is not None else None, self._wide, self._cap, C.byref(n_wide))
            if rc == F.MRQ_OK:
                break
            if n_wide.value > self._cap:  # the escape list was too small: the base is unchanged, retry with room
                self._cap = n_wide.value + n_wide.value // 4  # (kept for the next frames)
                self._wide = (F.Msg * self._cap)()
                continue
            raise ValueError((self.L.mrq_last_error(None) or b"mrq_pack8 failed").decode())
        msgs = [(int(m.group), int(m.from_), int(m.type), int(m.term), int(m.index), int(m.logterm), int(m.commit))
                for m in self._wide[: n_wide.value]]
        return word, p8, msgs


def unpack8(word: np.ndarray, self_id: np.ndarray, base_index: np.ndarray, base_term: np.ndarray, n_replicas: int):
    """`mrq_unpack8`: the device's decode of one byte-form frame, on the host.  -> (wide dense columns, slid base)."""
    L = F.load()
    G = int(self_id.shape[0])
    out = {"type": np.zeros((n_replicas, G), np.uint8)}
    for k in ("term", "index", "logterm", "commit"):
        out[k] = np.zeros((n_replicas, G), np.uint64)
    base = np.ascontiguousarray(base_index, np.uint64).copy()
    bt = np.ascontiguousarray(base_term, np.uint64)
    sid = np.ascontiguousarray(self_id, np.uint8)
    w = np.ascontiguousarray(word, np.uint8)
    view = F.InboxOut(_p(out["type"], F.u8p), _p(out["term"], F.u64p), _p(out["index"], F.u64p), _p(out["logterm"], F.u64p),
                      _p(out["commit"], F.u64p), None)
    rc = L.mrq_unpack8(_p(w, F.u8p) if w.size else None, _p(sid, F.u8p), G, n_replicas, _p(base, F.u64p), _p(bt, F.u64p), C.byref(view))
    if rc != F.MRQ_OK:
        raise ValueError((L.mrq_last_error(None) or b"mrq_unpack8 failed").decode())
    return out, base


class PinnedArray:
    """A numpy array over page-locked host memory from mrq_alloc_pinned (asynchronous H2D / D2H)."""

    def __init__(self, shape, dtype):
        self.L = F.load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.L.mrq_alloc_pinned(self.nbytes)
        if not self.ptr:
            raise MemoryError("mrq_alloc_pinned failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.L.mrq_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
