"""raftsql_b200 — Blackwell-native multi-raft quorum engine behind raftsql's raftPipe seam.

The product is `libmrq.so` (hand-written sm_100a CUDA behind the C-ABI in include/mrq.h); this package
is the Python host side of that ABI (ctypes) plus the host shim mirroring the reference's
`raftpipe.go` interface.  Importing the package does not load the library; constructing an Engine does,
and fails loudly if the library has not been built — there is no CPU fallback.
"""
from ._ffi import (MrqLibraryMissing, LIB_PATH, MSG_APP, MSG_APP_RESP, MSG_HEARTBEAT, MSG_HEARTBEAT_RESP, MSG_NONE,
                   MSG_REJECT, MSG_VOTE, MSG_VOTE_RESP, ROLE_CANDIDATE, ROLE_FOLLOWER, ROLE_LEADER, TraceParams)
from .engine import Engine, MrqError, empty_inbox, empty_state, preset_trace

__all__ = ["Engine", "MrqError", "MrqLibraryMissing", "TraceParams", "preset_trace", "empty_inbox", "empty_state",
           "LIB_PATH"]
