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
