"""More of upstream's known-answer tables (etcd raft/raft_test.go, raft/raft_paper_test.go, v2.2-v2.3 era, recalled;
every row is re-derived in its comment from the rule it exercises), each run against BOTH restatements — the C
oracle (oracle/raft_oracle.c) and the independent Python model (tests/pyraft_model.py).  The upstream source is
not available here (parity unpinned): two independent transcriptions agreeing with each other AND
with upstream's own expected values is the pin that is available."""
import pytest

from oracle import Oracle
from pyraft_model import Raft

MsgHup, MsgBeat, MsgProp, MsgApp, MsgAppResp, MsgVote, MsgVoteResp, MsgHeartbeat = 0, 1, 2, 3, 4, 5, 6, 8
FOLLOWER, CANDIDATE, LEADER = 0, 1, 2
OUT_CAMPAIGN, OUT_BECAME_LEADER, OUT_BCAST_APPEND, OUT_BCAST_HEARTBEAT, OUT_STEPPED_DOWN = 1, 2, 4, 8, 16


class OracleSM:
    """one raft state machine of the C oracle behind the few verbs the tables need"""

    def __init__(self, size, nid=1, log=()):
        self.o = Oracle(1, size, self_id=nid)
        if log:
            self.o.set_log(0, list(log))

    def step(self, type, frm=0, term=0, index=0, logterm=0, commit=0, reject=False, n=0):
        self.o.step(0, type, frm=frm, term=term, index=index, logterm=logterm, commit=commit, reject=reject, n_entries=n)

    def tick(self):
        self.o.tick(None)

    def clear_out(self):
        self.o.clear_out(0)

    def get(self, k):
        v = self.o.export()[k]
        return [int(x) for x in v[:, 0]] if v.ndim == 2 else int(v[0])


class PyModelSM:
    """the same verbs over tests/pyraft_model.py"""

    def __init__(self, size, nid=1, log=()):
        self.tick_no = 0
        self.r = Raft(nid, range(1, size + 1), 10, 1, 0x5EED + nid, 0, lambda: self.tick_no)
        self.r.raftLog.terms = list(log)
        self.size = size

    def step(self, type, frm=0, term=0, index=0, logterm=0, commit=0, reject=False, n=0):
        self.r.Step({"type": type, "from": frm or self.r.id, "term": term, "index": index, "logterm": logterm,
                     "commit": commit, "reject": reject, "n": n})

    def tick(self):
        self.r.out = 0
        self.r.tick()
        self.tick_no += 1

    def clear_out(self):
        self.r.out = 0

    def get(self, k):
        r = self.r
        return {"role": r.state, "term": r.Term, "vote": r.Vote, "lead": r.lead, "committed": r.raftLog.committed,
                "last_index": r.raftLog.lastIndex(), "last_term": r.raftLog.lastTerm(), "out": r.out,
                "match": [r.prs[p].Match for p in range(1, self.size + 1)],
                "votes": [0 if p not in r.votes else (1 if r.votes[p] else 2) for p in range(1, self.size + 1)],
                "election_elapsed": r.electionElapsed, "randomized_timeout": r.randomizedElectionTimeout}[k]


SMS = [pytest.param(OracleSM, id="c-oracle"), pytest.param(PyModelSM, id="py-model")]


def elect(sm, size):
    """becomeCandidate(); becomeLeader(): campaign, then as many grants as a quorum needs"""
    sm.step(MsgHup)
    term = sm.get("term")
    for p in range(2, size // 2 + 2):
        if sm.get("role") != LEADER:
            sm.step(MsgVoteResp, frm=p, term=term)
    assert sm.get("role") == LEADER
    return term


# raft_paper_test.go TestLeaderAcknowledgeCommit: (size, acceptors) -> the proposal is committed?
# the leader commits once a majority (itself included) has the entry: q = size/2 + 1
ACK_COMMIT = [(1, [], True), (3, [], False), (3, [2], True), (3, [2, 3], True), (5, [], False), (5, [2], False),
              (5, [2, 3], True), (5, [2, 3, 4], True), (5, [2, 3, 4, 5], True)]


@pytest.mark.parametrize("SM", SMS)
@pytest.mark.parametrize("size,acceptors,wack", ACK_COMMIT)
def test_TestLeaderAcknowledgeCommit(SM, size, acceptors, wack):
    sm = SM(size)
    term = elect(sm, size)
    # commitNoopEntry: every follower acknowledges the leader's empty entry
    li = sm.get("last_index")
    for p in range(2, size + 1):
        sm.step(MsgAppResp, frm=p, term=term, index=li)
    assert sm.get("committed") == li
    sm.step(MsgProp, n=1)
    li = sm.get("last_index")
    for p in acceptors:
        sm.step(MsgAppResp, frm=p, term=term, index=li)
    assert (sm.get("committed") >= li) == wack


# raft_test.go TestAllServerStepdown: any server that sees a higher term in a MsgVote / MsgApp becomes a follower
# of that term; its log is untouched (the MsgApp carries LogTerm 3 at Index 0: no match -> rejected);
# lead = None for MsgVote, the sender for MsgApp.           (state, wterm, windex)
@pytest.mark.parametrize("SM", SMS)
@pytest.mark.parametrize("mtype", [MsgVote, MsgApp])
@pytest.mark.parametrize("state,windex", [(FOLLOWER, 0), (CANDIDATE, 0), (LEADER, 1)])
def test_TestAllServerStepdown(SM, mtype, state, windex):
    sm = SM(3)
    if state == CANDIDATE:
        sm.step(MsgHup)
    elif state == LEADER:
        elect(sm, 3)
    assert sm.get("role") == state
    # host-resolved MsgApp (include/mrq.h): the append did not match, so REJECT and no log fields
    sm.step(mtype, frm=2, term=3, logterm=3, reject=(mtype == MsgApp))
    assert sm.get("role") == FOLLOWER
    assert sm.get("term") == 3
    assert sm.get("last_index") == windex
    assert sm.get("lead") == (2 if mtype == MsgApp else 0)


# raft_paper_test.go TestCandidateFallback: a candidate that receives an AppendEntries from a leader whose term is at
# least its own recognises it and returns to follower state, adopting that term
@pytest.mark.parametrize("SM", SMS)
@pytest.mark.parametrize("mterm", [1, 2])
def test_TestCandidateFallback(SM, mterm):
    sm = SM(3)
    sm.step(MsgHup)
    assert sm.get("role") == CANDIDATE and sm.get("term") == 1
    sm.step(MsgApp, frm=2, term=mterm, index=0, logterm=0, commit=0)
    assert sm.get("role") == FOLLOWER and sm.get("term") == mterm and sm.get("lead") == 2


# raft_test.go TestRecvMsgBeat / raft_paper_test.go TestLeaderBcastBeat: only a leader turns MsgBeat into heartbeats
@pytest.mark.parametrize("SM", SMS)
@pytest.mark.parametrize("state,wbeat", [(LEADER, True), (CANDIDATE, False), (FOLLOWER, False)])
def test_TestRecvMsgBeat(SM, state, wbeat):
    sm = SM(3)
    if state == CANDIDATE:
        sm.step(MsgHup)
    elif state == LEADER:
        elect(sm, 3)
    sm.clear_out()
    sm.step(MsgBeat)
    assert bool(sm.get("out") & OUT_BCAST_HEARTBEAT) == wbeat
    # and the leader's own clock does it every heartbeatTimeout (= 1) ticks
    sm.clear_out()
    sm.tick()
    assert bool(sm.get("out") & OUT_BCAST_HEARTBEAT) == wbeat


# raft_paper_test.go TestFollowerStartElection / TestCandidateStartNewElection: on election timeout a non-leader
# increments its term, becomes candidate, votes for itself and asks everybody else (the CAMPAIGN flag here)
@pytest.mark.parametrize("SM", SMS)
@pytest.mark.parametrize("state", [FOLLOWER, CANDIDATE])
def test_TestNonleaderStartElection(SM, state):
    sm = SM(3)
    term0 = 1
    if state == CANDIDATE:
        sm.step(MsgHup)  # term 1, candidate
    else:
        sm.step(MsgHeartbeat, frm=2, term=1)  # becomeFollower(1, 2)
    rto = sm.get("randomized_timeout")
    assert 10 <= rto < 20  # [electiontimeout, 2 * electiontimeout - 1]
    for _ in range(rto - sm.get("election_elapsed") - 1):
        sm.tick()
        assert sm.get("term") == term0 and sm.get("role") == state
    sm.clear_out()
    sm.tick()  # the timeout fires: Step(MsgHup)
    assert sm.get("term") == term0 + 1
    assert sm.get("role") == CANDIDATE
    assert sm.get("vote") == 1 and sm.get("votes")[0] == 1
    assert sm.get("out") & OUT_CAMPAIGN


# raft_paper_test.go TestLeaderCommitPrecedingEntries: when the leader commits an entry of its own term, every
# preceding entry commits with it.  Prior logs (entry terms): the new leader is at term 3.
@pytest.mark.parametrize("SM", SMS)
@pytest.mark.parametrize("log", [[], [2], [1, 2], [1]])
def test_TestLeaderCommitPrecedingEntries(SM, log):
    sm = SM(3, log=log)
    sm.step(MsgHeartbeat, frm=2, term=2)  # r.loadState(HardState{Term: 2})
    term = elect(sm, 3)
    assert term == 3
    sm.step(MsgProp, n=1)
    li = sm.get("last_index")
    assert li == len(log) + 2  # the empty entry of term 3, then the proposal
    assert sm.get("committed") == 0
    sm.step(MsgAppResp, frm=2, term=term, index=li)  # acceptAndReply from one follower = a quorum of 3
    assert sm.get("committed") == li and sm.get("last_term") == 3


# raft_test.go TestProposal, the rows about who accepts proposals: a leader appends, a candidate drops
# (and a follower without a leader drops; with one it forwards)
@pytest.mark.parametrize("SM", SMS)
def test_proposals_by_role(SM):
    OUT_PROP_DROPPED, OUT_PROP_FORWARD = 32, 64
    sm = SM(3)
    sm.clear_out()
    sm.step(MsgProp, n=1)  # follower, no leader known
    assert sm.get("out") & OUT_PROP_DROPPED and sm.get("last_index") == 0
    sm.step(MsgHeartbeat, frm=2, term=1)
    sm.clear_out()
    sm.step(MsgProp, n=1)  # follower of 2: forwarded
    assert sm.get("out") & OUT_PROP_FORWARD and sm.get("last_index") == 0
    sm.step(MsgHup)
    sm.clear_out()
    sm.step(MsgProp, n=1)  # candidate: dropped
    assert sm.get("out") & OUT_PROP_DROPPED and sm.get("last_index") == 0
    sm.step(MsgVoteResp, frm=2, term=sm.get("term"))
    li = sm.get("last_index")
    sm.step(MsgProp, n=2)  # leader: appended with its term
    assert sm.get("last_index") == li + 2 and sm.get("last_term") == sm.get("term")
    assert sm.get("match")[0] == li + 2
