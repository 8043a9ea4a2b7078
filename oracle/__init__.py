"""ctypes loader for the CPU oracle (oracle/raft_oracle.c).

TEST INFRASTRUCTURE ONLY — PARITY UNPINNED (see oracle/raft_oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under raftsql_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class TraceParams(C.Structure):
    """Mirror of include/mrq_trace.h `mrq_trace_params`."""

    _fields_ = [
        ("seed", C.c_uint64),
        ("p_ack_256", C.c_uint32),
        ("p_grant_256", C.c_uint32),
        ("p_reject_256", C.c_uint32),
        ("p_heartbeat_256", C.c_uint32),
        ("churn_65536", C.c_uint32),
        ("lagging_pct", C.c_uint32),
        ("max_prop", C.c_uint32),
        ("lag_kind", C.c_uint32),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "raft_oracle.c")
    deps = [src, os.path.join(_HERE, "raft_oracle.h"), os.path.join(_HERE, "..", "include", "mrq.h"),
            os.path.join(_HERE, "..", "include", "mrq_trace.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(d) > os.path.getmtime(_LIB_PATH) for d in deps if os.path.exists(d))
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_tick.argtypes = [C.c_void_p, u8p, u64p, u64p, u64p, u64p, u32p, C.c_int]
    L.orc_quorum_commit.argtypes = [C.c_void_p, C.c_int]
    L.orc_gen_trace.argtypes = [C.c_void_p, C.POINTER(TraceParams), C.c_uint64, u8p, u64p, u64p, u64p, u64p, u32p,
                                C.c_int]
    L.orc_export.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p, u64p, u64p, u64p, u8p, u8p, u8p, u8p, u16p, u16p,
                             u16p, u32p]
    L.orc_import.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p, u64p, u64p, u64p, u8p, u8p, u8p, u8p, u16p, u16p,
                             u16p]
    L.orc_step.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                           C.c_uint64, C.c_int, C.c_uint32]
    L.orc_group_set_log.argtypes = [C.c_void_p, C.c_uint64, u64p, C.c_size_t]
    L.orc_group_clear_out.argtypes = [C.c_void_p, C.c_uint64]
    L.orc_tick_count.restype = C.c_uint64
    L.orc_tick_count.argtypes = [C.c_void_p]
    L.orc_set_tick_count.argtypes = [C.c_void_p, C.c_uint64]
    L.orc_errors.restype = C.c_uint64
    L.orc_errors.argtypes = [C.c_void_p]
    L.orc_hw_threads.restype = C.c_int
    L.orc_kat_commit.restype = C.c_uint64
    L.orc_kat_commit.argtypes = [u64p, C.c_int, u64p, C.c_size_t, C.c_uint64]
    L.orc_quorum_index_bruteforce.restype = C.c_uint64
    L.orc_quorum_index_bruteforce.argtypes = [u64p, C.c_int]
    L.orc_raft_new.restype = C.c_void_p
    L.orc_raft_new.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int]
    L.orc_raft_free.argtypes = [C.c_void_p]
    L.orc_raft_set_log.argtypes = [C.c_void_p, u64p, C.c_size_t]
    L.raftLog_isUpToDate.restype = C.c_int
    _lib = L
    return L


def _p(a, ty):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ty)


STATE_COLUMNS = ("term", "vote", "committed", "last_index", "last_term", "term_start", "match", "role", "lead",
                 "self_id", "votes", "election_elapsed", "heartbeat_elapsed", "randomized_timeout")


def empty_state(G: int, R: int) -> dict:
    """Host arrays in the SoA layout shared by the oracle and the engine (include/mrq.h mrq_state)."""
    z64 = lambda *s: np.zeros(s, dtype=np.uint64)
    return dict(term=z64(G), vote=z64(G), committed=z64(G), last_index=z64(G), last_term=z64(G),
                term_start=z64(G), match=z64(R, G), role=np.zeros(G, np.uint8), lead=np.zeros(G, np.uint8),
                self_id=np.zeros(G, np.uint8), votes=np.zeros((R, G), np.uint8),
                election_elapsed=np.zeros(G, np.uint16), heartbeat_elapsed=np.zeros(G, np.uint16),
                randomized_timeout=np.zeros(G, np.uint16))


def empty_inbox(G: int, R: int) -> dict:
    return dict(type=np.zeros((R, G), np.uint8), term=np.zeros((R, G), np.uint64), index=np.zeros((R, G), np.uint64),
                logterm=np.zeros((R, G), np.uint64), commit=np.zeros((R, G), np.uint64),
                prop_count=np.zeros(G, np.uint32))


class Oracle:
    """G raft groups stepped one message at a time on the CPU."""

    def __init__(self, G: int, R: int, *, group_base: int = 0, election_tick: int = 10, heartbeat_tick: int = 1,
                 seed: int = 0, self_id: int = 0):
        self.L = lib()
        self.G, self.R = int(G), int(R)
        self.h = self.L.orc_create(G, R, group_base, election_tick, heartbeat_tick, seed, self_id)
        if not self.h:
            raise ValueError("orc_create rejected the configuration")

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tick(self, inbox: dict | None = None, nthreads: int = 1):
        ib = inbox or {}
        g = lambda k, ty: _p(ib.get(k), ty)
        self.L.orc_tick(self.h, g("type", u8p), g("term", u64p), g("index", u64p), g("logterm", u64p),
                        g("commit", u64p), g("prop_count", u32p), nthreads)

    def step(self, g: int, type: int, frm: int = 0, term: int = 0, index: int = 0, logterm: int = 0,
             commit: int = 0, reject: bool = False, n_entries: int = 0):
        """raft.Step(m) for one message on group g."""
        self.L.orc_step(self.h, g, type, frm, term, index, logterm, commit, int(reject), n_entries)

    def set_log(self, g: int, entry_terms):
        t = np.asarray(entry_terms, dtype=np.uint64)
        self.L.orc_group_set_log(self.h, g, _p(t, u64p), len(t))

    def clear_out(self, g: int):
        self.L.orc_group_clear_out(self.h, g)

    def quorum_commit(self, nthreads: int = 1):
        self.L.orc_quorum_commit(self.h, nthreads)

    def gen_trace(self, params: TraceParams, tick: int, nthreads: int = 1) -> dict:
        ib = empty_inbox(self.G, self.R)
        self.L.orc_gen_trace(self.h, C.byref(params), tick, _p(ib["type"], u8p), _p(ib["term"], u64p),
                             _p(ib["index"], u64p), _p(ib["logterm"], u64p), _p(ib["commit"], u64p),
                             _p(ib["prop_count"], u32p), nthreads)
        return ib

    def export(self) -> dict:
        s = empty_state(self.G, self.R)
        out = np.zeros(self.G, np.uint32)
        self.L.orc_export(self.h, _p(s["term"], u64p), _p(s["vote"], u64p), _p(s["committed"], u64p),
                          _p(s["last_index"], u64p), _p(s["last_term"], u64p), _p(s["term_start"], u64p),
                          _p(s["match"], u64p), _p(s["role"], u8p), _p(s["lead"], u8p), _p(s["self_id"], u8p),
                          _p(s["votes"], u8p), _p(s["election_elapsed"], u16p), _p(s["heartbeat_elapsed"], u16p),
                          _p(s["randomized_timeout"], u16p), _p(out, u32p))
        s["out"] = out
        return s

    def import_state(self, s: dict):
        g = lambda k, ty: _p(s.get(k), ty)
        self.L.orc_import(self.h, g("term", u64p), g("vote", u64p), g("committed", u64p), g("last_index", u64p),
                          g("last_term", u64p), g("term_start", u64p), g("match", u64p), g("role", u8p),
                          g("lead", u8p), g("self_id", u8p), g("votes", u8p), g("election_elapsed", u16p),
                          g("heartbeat_elapsed", u16p), g("randomized_timeout", u16p))

    @property
    def tick_count(self) -> int:
        return int(self.L.orc_tick_count(self.h))

    @tick_count.setter
    def tick_count(self, t: int):
        self.L.orc_set_tick_count(self.h, t)

    @property
    def errors(self) -> int:
        return int(self.L.orc_errors(self.h))


def kat_commit(matches, entry_terms, sm_term) -> int:
    m = np.asarray(matches, dtype=np.uint64)
    t = np.asarray(entry_terms, dtype=np.uint64)
    return int(lib().orc_kat_commit(_p(m, u64p), len(m), _p(t, u64p), len(t), sm_term))


def quorum_index_bruteforce(matches) -> int:
    m = np.ascontiguousarray(matches, dtype=np.uint64)
    return int(lib().orc_quorum_index_bruteforce(_p(m, u64p), len(m)))


def hw_threads() -> int:
    return int(lib().orc_hw_threads())
