"""Upstream's known-answer tables of tests/test_oracle_kat2.py (ack-commit, stepdown, candidate fallback, start of
an election, preceding entries, proposals by role) replayed THROUGH THE GPU ENGINE, one message per tick: the same
test bodies, with the engine behind the state-machine verbs instead of a CPU restatement.
The bodies are also rehearsed on the CPU against tests/engine_double.py (tests/test_rehearsals_cpu.py)."""
import numpy as np
import pytest

import test_oracle_kat2 as k
from raftsql_b200 import Engine
from raftsql_b200 import _ffi as F

pytestmark = [pytest.mark.gpu]


class EngineSM:
    """one raft state machine = an engine of one group; a Step is one tick with that single message in the inbox
    (ElectionTick is huge, so the Tick() that follows every Step never fires a timer on its own)"""

    ET = 1000

    def __init__(self, size, nid=1, log=()):
        self.size = size
        self.e = Engine(1, size, self_id=nid, election_tick=self.ET, seed=7)
        if log:
            self.e.import_state({"last_index": np.array([len(log)], np.uint64), "last_term": np.array([log[-1]], np.uint64)})

    def step(self, type, frm=0, term=0, index=0, logterm=0, commit=0, reject=False, n=0):
        if type == k.MsgHup:  # campaign(): make this tick the election timeout
            self.e.import_state({"randomized_timeout": np.array([1], np.uint16), "election_elapsed": np.array([0], np.uint16)})
            self.e.tick_idle(1)
        elif type == k.MsgProp:
            self.e.clear_inbox(0)
            self.e.propose([0], [n], slot=0)
            self.e.tick(0)
        elif type == k.MsgBeat:
            pytest.skip("MsgBeat is a local message of the node's own clock: the engine raises it in Tick(), not from the inbox")
        else:
            self.e.post_inbox_delta([(0, frm, type | (F.MSG_REJECT if reject else 0), term, index, logterm, commit)], slot=0)
            self.e.tick(0)

    def tick(self):
        self.e.tick_idle(1)

    def clear_out(self):
        pass  # the out word is per tick

    def get(self, key):
        if key == "out":
            return int(self.e.sync_out()[0])
        v = self.e.export_state()[key]
        return [int(x) for x in v[:, 0]] if v.ndim == 2 else int(v[0])


@pytest.mark.parametrize("size,acceptors,wack", k.ACK_COMMIT)
def test_TestLeaderAcknowledgeCommit_engine(size, acceptors, wack):
    k.test_TestLeaderAcknowledgeCommit(EngineSM, size, acceptors, wack)


@pytest.mark.parametrize("mtype", [k.MsgVote, k.MsgApp])
@pytest.mark.parametrize("state,windex", [(k.FOLLOWER, 0), (k.CANDIDATE, 0), (k.LEADER, 1)])
def test_TestAllServerStepdown_engine(mtype, state, windex):
    k.test_TestAllServerStepdown(EngineSM, mtype, state, windex)


@pytest.mark.parametrize("mterm", [1, 2])
def test_TestCandidateFallback_engine(mterm):
    k.test_TestCandidateFallback(EngineSM, mterm)


@pytest.mark.parametrize("log", [[], [2], [1, 2], [1]])
def test_TestLeaderCommitPrecedingEntries_engine(log):
    k.test_TestLeaderCommitPrecedingEntries(EngineSM, log)


def test_proposals_by_role_engine():
    k.test_proposals_by_role(EngineSM)


def test_only_a_leaders_clock_raises_heartbeats_engine():
    """TestLeaderBcastBeat, the half the engine owns: every heartbeatTimeout (= 1) ticks a leader broadcasts, a
    follower or candidate never does"""
    for state in (k.FOLLOWER, k.CANDIDATE, k.LEADER):
        sm = EngineSM(3)
        if state == k.CANDIDATE:
            sm.step(k.MsgHup)
        elif state == k.LEADER:
            k.elect(sm, 3)
        assert sm.get("role") == state
        sm.tick()
        assert bool(sm.get("out") & k.OUT_BCAST_HEARTBEAT) == (state == k.LEADER)
