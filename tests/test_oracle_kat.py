"""Pin the CPU oracle against etcd-raft's own known-answer tables.

The tables are the ones SURVEY.md §8c records as [UPSTREAM-RECALLED] from etcd raft/raft_test.go and
raft/raft_paper_test.go (v2.2–v2.3 era): the upstream source is not in /root/reference (the reference
imports it, raft.go:30) and is not on this box, so every row was also re-derived by hand from the rules
in SURVEY §8a rows a7–a16 before being committed here.
"""
import numpy as np
import pytest

import oracle
from oracle import Oracle

MsgHup, MsgProp, MsgApp, MsgAppResp, MsgVote, MsgVoteResp, MsgHeartbeat = 0, 2, 3, 4, 5, 6, 8
FOLLOWER, CANDIDATE, LEADER = 0, 1, 2

OUT_CAMPAIGN = 0x01
OUT_VOTE_REPLY_SHIFT = 8

# upstream raft_test.go TestCommit: (matches, log entry terms, smTerm) -> committed
TEST_COMMIT = [
    # single
    ([1], [1], 1, 1),
    ([1], [1], 2, 0),
    ([2], [1, 2], 2, 2),
    ([1], [2], 2, 1),
    # odd
    ([2, 1, 1], [1, 2], 1, 1),
    ([2, 1, 1], [1, 1], 2, 0),
    ([2, 1, 2], [1, 2], 2, 2),
    ([2, 1, 2], [1, 1], 2, 0),
    # even
    ([2, 1, 1, 1], [1, 2], 1, 1),
    ([2, 1, 1, 1], [1, 1], 2, 0),
    ([2, 1, 1, 2], [1, 2], 1, 1),
    ([2, 1, 1, 2], [1, 1], 2, 0),
    ([2, 1, 2, 2], [1, 2], 2, 2),
    ([2, 1, 2, 2], [1, 1], 2, 0),
]


@pytest.mark.parametrize("matches,terms,sm_term,want", TEST_COMMIT)
def test_commit_kat(matches, terms, sm_term, want):
    assert oracle.kat_commit(matches, terms, sm_term) == want


# upstream raft_paper_test.go TestVoter: (voter log terms, cand logterm, cand index) -> reject?
TEST_VOTER = [
    ([1], 1, 1, False),
    ([1], 1, 2, False),
    ([1, 1], 1, 1, True),
    ([1], 2, 1, False),
    ([1], 2, 2, False),
    ([1, 1], 2, 1, False),
    ([2], 1, 1, True),
    ([2], 1, 2, True),
    ([2, 1], 1, 1, True),
]


def _vote_reply(out_word, frm):
    return (int(out_word) >> (OUT_VOTE_REPLY_SHIFT + 2 * (frm - 1))) & 3


@pytest.mark.parametrize("log,logterm,index,wreject", TEST_VOTER)
def test_voter_kat(log, logterm, index, wreject):
    o = Oracle(1, 2, self_id=1)
    o.set_log(0, log)
    # r.Step(pb.Message{From: 2, To: 1, Type: MsgVote, Term: 3, LogTerm: logterm, Index: index})
    o.step(0, MsgVote, frm=2, term=3, index=index, logterm=logterm)
    s = o.export()
    assert _vote_reply(s["out"][0], 2) == (2 if wreject else 1)
    assert s["term"][0] == 3
    assert s["vote"][0] == (0 if wreject else 2)


# upstream raft_paper_test.go TestFollowerVote: (vote, nvote) -> reject?
@pytest.mark.parametrize("vote,nvote,wreject", [(0, 1, False), (0, 2, False), (1, 1, False), (2, 2, False),
                                                (1, 2, True), (2, 1, True)])
def test_follower_vote_kat(vote, nvote, wreject):
    o = Oracle(1, 3, self_id=3)
    st = o.export()
    st["term"][0] = 1
    st["vote"][0] = vote
    o.import_state({"term": st["term"], "vote": st["vote"]})
    o.step(0, MsgVote, frm=nvote, term=1, index=0, logterm=0)
    s = o.export()
    assert _vote_reply(s["out"][0], nvote) == (2 if wreject else 1)


# upstream raft_paper_test.go TestLeaderElectionInOneRoundRPC: (size, votes{id: granted}) -> state
ONE_ROUND = [
    (1, {}, LEADER),
    (3, {2: True, 3: True}, LEADER),
    (3, {2: True}, LEADER),
    (5, {2: True, 3: True, 4: True, 5: True}, LEADER),
    (5, {2: True, 3: True, 4: True}, LEADER),
    (5, {2: True, 3: True}, LEADER),
    (3, {2: False, 3: False}, FOLLOWER),
    (5, {2: False, 3: False, 4: False, 5: False}, FOLLOWER),
    (5, {2: True, 3: False, 4: False, 5: False}, FOLLOWER),
    (3, {}, CANDIDATE),
    (5, {2: True}, CANDIDATE),
    (5, {2: False, 3: False}, CANDIDATE),
    (5, {}, CANDIDATE),
]


@pytest.mark.parametrize("size,votes,want", ONE_ROUND)
def test_election_one_round_kat(size, votes, want):
    o = Oracle(1, size, self_id=1)
    o.step(0, MsgHup, frm=1)
    for vid, granted in votes.items():
        o.step(0, MsgVoteResp, frm=vid, term=1, reject=not granted)
    s = o.export()
    assert s["role"][0] == want
    assert s["term"][0] == 1


class Network:
    """upstream raft_test.go `network`: n rafts (None = nopStepper / dead) exchanging the messages that the
    oracle's out word stands for (CAMPAIGN -> MsgVote to all, vote-reply bits -> MsgVoteResp)."""

    def __init__(self, logs):
        self.n = len(logs)
        self.nodes = []
        for i, log in enumerate(logs):
            if log == "dead":
                self.nodes.append(None)
                continue
            o = Oracle(1, self.n, self_id=i + 1)
            if log:
                o.set_log(0, log)
            self.nodes.append(o)

    def campaign(self, nid):
        cand = self.nodes[nid - 1]
        cand.clear_out(0)
        cand.step(0, MsgHup, frm=nid)
        s = cand.export()
        if not (int(s["out"][0]) & OUT_CAMPAIGN):
            return
        replies = []
        for pid in range(1, self.n + 1):
            peer = self.nodes[pid - 1]
            if pid == nid or peer is None:
                continue
            peer.clear_out(0)
            peer.step(0, MsgVote, frm=nid, term=int(s["term"][0]), index=int(s["last_index"][0]),
                      logterm=int(s["last_term"][0]))
            ps = peer.export()
            rep = _vote_reply(ps["out"][0], nid)
            if rep:
                replies.append((pid, int(ps["term"][0]), rep == 2))
        for pid, term, reject in replies:
            cand.step(0, MsgVoteResp, frm=pid, term=term, reject=reject)


# upstream raft_test.go TestLeaderElection
LEADER_ELECTION = [
    ([None, None, None], LEADER),
    ([None, None, "dead"], LEADER),
    ([None, "dead", "dead"], CANDIDATE),
    ([None, "dead", "dead", None], CANDIDATE),
    ([None, "dead", "dead", None, None], LEADER),
    # three logs further along than 0
    ([None, [1], [2], [1, 3], None], FOLLOWER),
    # logs converge
    ([[1], None, [2], [1], None], LEADER),
]


@pytest.mark.parametrize("logs,want", LEADER_ELECTION)
def test_leader_election_kat(logs, want):
    nt = Network(logs)
    nt.campaign(1)
    s = nt.nodes[0].export()
    assert s["role"][0] == want
    assert s["term"][0] == 1


def test_become_leader_appends_empty_entry_and_gates_commit():
    """becomeLeader appends one entry at the new term; earlier entries cannot be committed by counting
    replicas until an entry of the leader's own term reaches a quorum (Raft §5.4.2; upstream
    raftLog.maybeCommit's term check)."""
    o = Oracle(1, 3, self_id=1)
    o.set_log(0, [1, 1, 1])  # three entries from an older term
    st = o.export()
    st["term"][0] = 1
    o.import_state({"term": st["term"]})
    o.set_log(0, [1, 1, 1])
    o.step(0, MsgHup, frm=1)
    o.step(0, MsgVoteResp, frm=2, term=2)
    s = o.export()
    assert s["role"][0] == LEADER and s["term"][0] == 2
    assert s["last_index"][0] == 4 and s["last_term"][0] == 2 and s["term_start"][0] == 4
    assert list(s["match"][:, 0]) == [4, 0, 0]
    # follower 2 acknowledges index 3 (old-term entries only): quorum index 3, but term(3) != 2
    o.step(0, MsgAppResp, frm=2, term=2, index=3)
    assert o.export()["committed"][0] == 0
    # follower 2 acknowledges the new leader's entry: everything up to 4 commits at once
    o.step(0, MsgAppResp, frm=2, term=2, index=4)
    assert o.export()["committed"][0] == 4


def test_step_term_rules():
    o = Oracle(1, 3, self_id=1)
    o.step(0, MsgHup, frm=1)
    o.step(0, MsgVoteResp, frm=2, term=1)
    assert o.export()["role"][0] == LEADER
    o.step(0, MsgVote, frm=3, term=1, index=9, logterm=9)  # same term: leader rejects
    s = o.export()
    assert _vote_reply(s["out"][0], 3) == 2 and s["role"][0] == LEADER
    # higher term MsgVote: becomeFollower(term, None), Vote reset, then the grant decision
    o.step(0, MsgVote, frm=3, term=5, index=9, logterm=9)
    s = o.export()
    assert s["role"][0] == FOLLOWER and s["term"][0] == 5 and s["lead"][0] == 0 and s["vote"][0] == 3
    # higher term heartbeat: follower of that leader
    o.step(0, MsgHeartbeat, frm=2, term=6, commit=0)
    s = o.export()
    assert s["term"][0] == 6 and s["lead"][0] == 2 and s["vote"][0] == 0
