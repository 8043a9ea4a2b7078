"""Host node + consensus core against upstream's own unit tables for the parts the host shares with the core
(SURVEY §8f rows f2/f4): etcd raft/raft_test.go TestHandleMsgApp, TestHandleHeartbeat, TestHandleHeartbeatResp and
TestLeaderAppResp (v2.2-v2.3 era, recalled — the upstream source is not available here, so every row is also
re-derived in a comment from the algorithm upstream documents: raftLog.maybeAppend / commitTo / stepLeader).

In the reference all of this happened inside `raft.Step`; here the log half lives in raftsql_b200.hostnode
(`Log.maybe_append`, `_resolve_append`, `sendAppend` bookkeeping) and the index arithmetic in the core, joined
by the host-resolved MsgApp contract of include/mrq.h.  The tests drive ONE node through its real entry point
(`step_tick`, with hand-made peer messages in the transport) and read the peers' mailboxes.
The core here is the CPU checker (this is the CPU suite); tests/test_plumbing.py runs the same host code over
the GPU engine under `-m gpu`."""
import numpy as np
import pytest

from oracle_core import make_oracle_core
from raftsql_b200 import _ffi as F
from raftsql_b200.hostnode import HostNode, LocalTransport, Message


def u64(x):
    return np.array([x], np.uint64)


def follower_with_log(terms, term, committed=0, lead=0, npeers=3):
    """sm := newRaft(1, [1,2,3]); storage.Append(ents); sm.becomeFollower(term, lead); commitTo(committed)"""
    tr = LocalTransport()
    for p in range(2, npeers + 1):
        tr.register(p)
    core = make_oracle_core(npeers, 1)
    node = HostNode(core, 1, npeers, tr)
    node.log.ents = [(t, b"x") for t in terms]
    node.term, node.commit, node.lead = term, committed, lead
    core.import_state({"term": u64(term), "vote": u64(0), "committed": u64(committed), "last_index": u64(len(terms)),
                       "last_term": u64(terms[-1] if terms else 0), "lead": np.array([lead], np.uint8)})
    return node, tr


# (index, logterm, commit, entries[(term)], windex, wcommit, wreject) — log is [1:t1, 2:t2], committed 0
HANDLE_MSGAPP = [
    # Ensure 1: reply false if log doesn't contain an entry at prevLogIndex whose term matches prevLogTerm
    (2, 3, 3, [], 2, 0, True),          # term(2) = 2 != 3
    (3, 3, 3, [], 2, 0, True),          # no entry 3
    # Ensure 2: conflicts truncate, new entries append
    (1, 1, 1, [], 2, 1, False),         # nothing to append; commit = min(1, lastnewi 1)
    (0, 0, 1, [2], 1, 1, False),        # entry 1 conflicts (t1 vs t2): log becomes [1:t2]
    (2, 2, 3, [2, 2], 4, 3, False),     # append 3,4; commit = min(3, 4)
    (2, 2, 4, [2], 3, 3, False),        # append 3;   commit = min(4, 3)
    (1, 1, 4, [2], 2, 2, False),        # entry 2 already there; commit = min(4, 2)
    # Ensure 3: commit = min(leaderCommit, index of last new entry)
    (1, 1, 3, [], 2, 1, False),
    (1, 1, 3, [2], 2, 2, False),
    (2, 2, 3, [], 2, 2, False),
    (2, 2, 4, [], 2, 2, False),
]


@pytest.mark.parametrize("index,logterm,commit,ents,windex,wcommit,wreject", HANDLE_MSGAPP)
def test_TestHandleMsgApp(index, logterm, commit, ents, windex, wcommit, wreject):
    node, tr = follower_with_log([1, 2], term=2)
    tr.send([Message(F.MSG_APP, 1, 2, term=2, logterm=logterm, index=index, commit=commit, entries=[(t, b"y") for t in ents])])
    node.step_tick()
    assert node.log.last_index() == windex
    assert node.commit == wcommit
    s = node.core.export_state()
    assert int(s["last_index"][0]) == windex and int(s["committed"][0]) == wcommit  # host log and core agree
    assert int(s["last_term"][0]) == node.log.last_term()
    replies = tr.drain(2)
    assert len(replies) == 1 and replies[0].type == F.MSG_APP_RESP and replies[0].term == 2
    assert replies[0].reject == wreject
    if wreject:  # handleAppendEntries: Index = m.Index, RejectHint = lastIndex
        assert replies[0].index == index and replies[0].reject_hint == 2
    else:        # Index = lastnewi
        assert replies[0].index == index + len(ents)
    assert node.lead == 2 and node.role == F.ROLE_FOLLOWER  # stepFollower: r.lead = m.From


def test_msgapp_below_the_commit_index_is_answered_with_the_commit_index():
    """handleAppendEntries: `if m.Index < r.raftLog.committed { send MsgAppResp{Index: committed}; return }`"""
    node, tr = follower_with_log([1, 2, 2], term=2, committed=2, lead=2)
    tr.send([Message(F.MSG_APP, 1, 2, term=2, logterm=1, index=1, commit=3, entries=[(9, b"bogus")])])
    node.step_tick()
    assert node.log.last_index() == 3 and [t for t, _ in node.log.ents] == [1, 2, 2]  # untouched
    assert node.commit == 2
    (rep,) = tr.drain(2)
    assert rep.type == F.MSG_APP_RESP and not rep.reject and rep.index == 2


@pytest.mark.parametrize("m_commit,wcommit", [(3, 3), (1, 2)])  # commit + 1 -> advances; commit - 1 -> never decreases
def test_TestHandleHeartbeat(m_commit, wcommit):
    node, tr = follower_with_log([1, 2, 3], term=2, committed=2, lead=2)
    tr.send([Message(F.MSG_HEARTBEAT, 1, 2, term=2, commit=m_commit)])
    node.step_tick()
    assert node.commit == wcommit
    (rep,) = tr.drain(2)
    assert rep.type == F.MSG_HEARTBEAT_RESP and rep.term == 2


def make_leader(terms):
    """newRaft(1,[1,2,3]) with a log, then becomeCandidate(); becomeLeader() by way of a real election: the
    node times out, campaigns, node 2 grants.  Returns the node as leader with its empty entry appended."""
    node, tr = follower_with_log(terms, term=max(terms) if terms else 0)
    for _ in range(40):
        node.step_tick()
        if any(m.type == F.MSG_VOTE for m in tr.boxes[2]):
            break
    votes = [m for m in tr.drain(2) if m.type == F.MSG_VOTE]
    tr.drain(3)
    assert votes and votes[0].index == len(terms) and votes[0].logterm == (terms[-1] if terms else 0)
    tr.send([Message(F.MSG_VOTE_RESP, 1, 2, term=node.term)])
    node.step_tick()
    assert node.role == F.ROLE_LEADER and node.log.last_index() == len(terms) + 1
    assert node.log.ents[-1] == (node.term, b"")  # becomeLeader appends an empty entry of the new term
    return node, tr


def test_TestHandleHeartbeatResp():
    """'a heartbeat response will re-send log entries if the follower is behind' — and stops once it has acked."""
    node, tr = make_leader([1, 2, 3])
    last = node.log.last_index()
    first = [m for m in tr.drain(2) if m.type == F.MSG_APP]
    assert first and first[0].index + len(first[0].entries) == last  # bcastAppend on winning
    for _ in range(2):  # every heartbeat response from a follower that is behind triggers a sendAppend from Match+1
        tr.send([Message(F.MSG_HEARTBEAT_RESP, 1, 2, term=node.term)])
        node.step_tick()
        apps = [m for m in tr.drain(2) if m.type == F.MSG_APP]
        assert len(apps) == 1 and apps[0].index == 0 and len(apps[0].entries) == last
        assert apps[0].logterm == 0 and apps[0].term == node.term
    tr.send([Message(F.MSG_APP_RESP, 1, 2, term=node.term, index=last)])  # the follower catches up
    node.step_tick()
    tr.drain(2)
    tr.send([Message(F.MSG_HEARTBEAT_RESP, 1, 2, term=node.term)])
    node.step_tick()
    assert not [m for m in tr.drain(2) if m.type == F.MSG_APP]  # nothing left to send: heartbeats only


# Upstream's TestLeaderAppResp, on a fresh leader of term T whose log is [1:t1, 2:t1] + its own empty entry 3:T.
# Its columns live in two places here — Match / committed in the core, Next and the MsgApp that follows in the host:
#   reject at index 2, hint 1            -> match 0, no commit, Next = min(2, hint + 1) = 2: a probe MsgApp{Index: 1}
#   ack 2 (an index of an older term)    -> match 2, commit stays 0 (only current-term entries commit by counting)
#   ack 3                                -> match 3, commit 3 (quorum {1, 2} on an entry of term T), followers told
#   an ack carrying an older term        -> dropped by the term rule
def test_TestLeaderAppResp():
    node, tr = make_leader([1, 1])
    T, last = node.term, node.log.last_index()
    assert last == 3
    tr.drain(2), tr.drain(3)

    def step(msg):
        tr.send([msg])
        node.step_tick()
        s = node.core.export_state()
        step.to2 = tr.drain(2)
        return int(s["match"][1, 0]), int(s["committed"][0]), [m for m in step.to2 if m.type == F.MSG_APP]

    # denied: the leader does not commit, decreases Next and probes (maybeDecrTo: Next = min(rejected, hint + 1))
    match, committed, apps = step(Message(F.MSG_APP_RESP, 1, 2, term=T, index=2, reject=True, reject_hint=1))
    assert (match, committed) == (0, 0)
    assert len(apps) == 1 and apps[0].index == 1 and apps[0].logterm == 1 and len(apps[0].entries) == 2  # the probe from Next = 2
    # an ack of an old-term index advances Match but never the commit index
    match, committed, _ = step(Message(F.MSG_APP_RESP, 1, 2, term=T, index=2))
    assert (match, committed) == (2, 0)
    # accepted up to the leader's own entry: quorum (1 and 2) on an entry of the current term -> commit, broadcast
    match, committed, apps = step(Message(F.MSG_APP_RESP, 1, 2, term=T, index=3))
    assert (match, committed) == (3, 3) and node.commit == 3
    # the followers learn it in the same tick.  (Upstream's bcastAppend would carry it in an empty MsgApp; this host
    # has nothing left to append to either peer, and with HeartbeatTick = 1 (raft.go:155) the heartbeat of the same
    # tick carries it under upstream's own rule, commit = min(pr.Match, committed).)
    hb2 = [m for m in step.to2 if m.type == F.MSG_HEARTBEAT]
    hb3 = [m for m in tr.drain(3) if m.type == F.MSG_HEARTBEAT]
    assert hb2 and hb2[-1].commit == 3 and hb2[-1].term == T
    assert hb3 and hb3[-1].commit == 0  # node 3 has acknowledged nothing: it must not be told to commit what it lacks
    # a stale (lower-term) ack is ignored by the term rule
    match, committed, _ = step(Message(F.MSG_APP_RESP, 1, 2, term=T - 1, index=9))
    assert (match, committed) == (3, 3)


def test_msgapp_behind_a_higher_term_message_of_the_same_tick_leaves_log_and_wal_alone(tmp_path):
    """the engine Steps a tick's messages in sender order.  Node 3 (term 5) hears, in ONE tick,
    MsgVote(term 7) from node 1 and MsgApp(term 5) from node 2: the engine moves to term 7 first and drops the append
    on the term rule — so the host must not have appended / truncated / saved anything for it."""
    tr = LocalTransport()
    for p in (1, 2):
        tr.register(p)
    core = make_oracle_core(3, 3)
    node = HostNode(core, 3, 3, tr, str(tmp_path / "raftsql-3"))
    node.start()
    node.log.ents = [(5, b"a"), (5, b"b")]
    node.term, node.commit, node.lead = 5, 0, 2
    core.import_state({"term": u64(5), "vote": u64(0), "committed": u64(0), "last_index": u64(2), "last_term": u64(5),
                       "lead": np.array([2], np.uint8)})
    wal_size = lambda: __import__("os").path.getsize(node.wal.path)
    before = wal_size()
    tr.send([Message(F.MSG_APP, 3, 2, term=5, logterm=5, index=1, commit=0, entries=[(5, b"conflict-free"), (5, b"more")]),
             Message(F.MSG_VOTE, 3, 1, term=7, logterm=9, index=9)])
    node.step_tick()
    assert node.term == 7 and node.vote == 1                      # the vote was granted at the new term
    assert [d for _, d in node.log.ents] == [b"a", b"b"]          # the stale append changed nothing
    s = node.core.export_state()
    assert (int(s["last_index"][0]), int(s["last_term"][0])) == (2, 5)
    from raftsql_b200.hostnode import scan_records
    recs = scan_records(node.wal.path)[0]
    assert all("e" not in r and "t" not in r for r in recs), recs  # only the hardstate of the new term was saved
    assert wal_size() > before and recs[-1]["hs"][:2] == [7, 1]
    assert not [m for m in tr.drain(2) if m.type == F.MSG_APP_RESP]  # and nothing was acknowledged to the old leader
    node.stop()


def test_two_msgapps_in_one_tick_are_resolved_one_per_tick():
    node, tr = follower_with_log([1, 2], term=2)
    tr.send([Message(F.MSG_APP, 1, 2, term=2, logterm=2, index=2, commit=0, entries=[(2, b"x3")]),
             Message(F.MSG_APP, 1, 3, term=2, logterm=2, index=2, commit=0, entries=[(2, b"other3")])])
    node.step_tick()
    assert node.log.last_index() == 3 and len(node.backlog) == 1   # the second append waits for the next tick
    node.step_tick()
    assert node.log.last_index() == 3 and not node.backlog
