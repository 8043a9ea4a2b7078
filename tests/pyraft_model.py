"""A second, independent restatement of etcd-raft's per-message rules (v2.2–v2.3 era), in Python and in upstream's own
shape — `prs` and `votes` are maps, the log is a list of entry terms, `maybeCommit` sorts a slice — written
separately from oracle/raft_oracle.c (arrays, run-length log, C).  tests/test_oracle_vs_pymodel.py steps both on the
same random traces: two independent transcriptions of the same published algorithm must agree bit for bit, which
is the strongest pin available while the upstream source itself is out of reach (parity unpinned).

Only what the engine models is restated (no snapshots, no conf changes, checkQuorum off, MsgApp host-resolved);
the stand-in for `r.rand` and the MRQ_OUT_* word are the shared conventions of include/mrq_trace.h / mrq.h.
"""
MASK = (1 << 64) - 1
None_ = 0
Follower, Candidate, Leader = 0, 1, 2
MsgHup, MsgBeat, MsgProp, MsgApp, MsgAppResp, MsgVote, MsgVoteResp, MsgHeartbeat, MsgHeartbeatResp = 0, 1, 2, 3, 4, 5, 6, 8, 9

OUT_CAMPAIGN, OUT_BECAME_LEADER, OUT_BCAST_APPEND, OUT_BCAST_HEARTBEAT = 1, 2, 4, 8
OUT_STEPPED_DOWN, OUT_PROP_DROPPED, OUT_PROP_FORWARD, OUT_COMMIT_ADVANCED = 16, 32, 64, 128


def mix64(x):
    x &= MASK
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & MASK
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & MASK
    x ^= x >> 31
    return x


def rand4(seed, a, b, c):
    x = mix64((seed + 0x9E3779B97F4A7C15 * (a + 1)) & MASK)
    x = mix64(x ^ ((0xD1B54A32D192ED03 * (b + 1)) & MASK))
    x = mix64(x ^ ((0x8CB92BA72F3D8DD7 * (c + 1)) & MASK))
    return x


def randomized_timeout(seed, gg, tick_no, election_tick):
    return election_tick + rand4(seed, gg, tick_no, 0x7133) % election_tick


class Progress:
    def __init__(self, match=0, next_=1):
        self.Match, self.Next = match, next_

    def maybeUpdate(self, n):
        updated = False
        if self.Match < n:
            self.Match = n
            updated = True
        if self.Next < n + 1:
            self.Next = n + 1
        return updated


class RaftLog:
    def __init__(self):
        self.terms = []  # terms[i-1] = term of entry i
        self.committed = 0

    def lastIndex(self):
        return len(self.terms)

    def term(self, i):
        return self.terms[i - 1] if 1 <= i <= len(self.terms) else 0

    def lastTerm(self):
        return self.term(self.lastIndex())

    def isUpToDate(self, lasti, term):
        return term > self.lastTerm() or (term == self.lastTerm() and lasti >= self.lastIndex())

    def commitTo(self, tocommit):
        if self.committed < tocommit:
            if self.lastIndex() < tocommit:
                return False  # upstream panics
            self.committed = tocommit
        return True

    def maybeCommit(self, maxIndex, term):
        if maxIndex > self.committed and self.term(maxIndex) == term:
            self.commitTo(maxIndex)
            return True
        return False


class Raft:
    def __init__(self, id_, peers, election_tick, heartbeat_tick, seed, gg, clock):
        self.id, self.Term, self.Vote, self.lead, self.state = id_, 0, None_, None_, Follower
        self.prs = {p: Progress() for p in peers}
        self.votes = {}
        self.raftLog = RaftLog()
        self.electionElapsed = self.heartbeatElapsed = 0
        self.electionTimeout, self.heartbeatTimeout = election_tick, heartbeat_tick
        self.seed, self.gg, self.clock = seed, gg, clock  # clock: callable -> current tick number
        self.randomizedElectionTimeout = 0
        self.out = 0
        self.errors = 0
        self.becomeFollower(0, None_)
        self.out = 0

    def q(self):
        return len(self.prs) // 2 + 1

    def reset(self, term):
        if self.Term != term:
            self.Term = term
            self.Vote = None_
        self.lead = None_
        self.electionElapsed = self.heartbeatElapsed = 0
        self.randomizedElectionTimeout = randomized_timeout(self.seed, self.gg, self.clock(), self.electionTimeout)
        self.votes = {}
        for p in self.prs:
            self.prs[p] = Progress(0, self.raftLog.lastIndex() + 1)
            if p == self.id:
                self.prs[p].Match = self.raftLog.lastIndex()

    def becomeFollower(self, term, lead):
        if self.state != Follower:
            self.out |= OUT_STEPPED_DOWN
        self.reset(term)
        self.lead, self.state = lead, Follower

    def becomeCandidate(self):
        self.reset(self.Term + 1)
        self.Vote, self.state = self.id, Candidate

    def becomeLeader(self):
        self.reset(self.Term)
        self.lead, self.state = self.id, Leader
        self.out |= OUT_BECAME_LEADER
        self.appendEntry(1)

    def appendEntry(self, n):
        self.raftLog.terms.extend([self.Term] * n)
        self.prs[self.id].maybeUpdate(self.raftLog.lastIndex())
        self.maybeCommit()

    def maybeCommit(self):
        mis = sorted((pr.Match for pr in self.prs.values()), reverse=True)
        mci = mis[self.q() - 1]
        ok = self.raftLog.maybeCommit(mci, self.Term)
        if ok:
            self.out |= OUT_COMMIT_ADVANCED
        return ok

    def poll(self, id_, v):
        if id_ not in self.votes:
            self.votes[id_] = v
        return sum(1 for vv in self.votes.values() if vv)

    def campaign(self):
        self.becomeCandidate()
        if self.q() == self.poll(self.id, True):
            self.becomeLeader()
            return
        self.out |= OUT_CAMPAIGN

    def _replyVote(self, to, reject):
        self.out |= (2 if reject else 1) << (8 + 2 * (to - 1))

    def _replyAck(self, to):
        self.out |= 1 << (24 + (to - 1))

    def _commitTo(self, c):
        before = self.raftLog.committed
        if not self.raftLog.commitTo(c):
            self.errors += 1
        if self.raftLog.committed != before:
            self.out |= OUT_COMMIT_ADVANCED

    def handleAppendEntries(self, m):  # host-resolved form (include/mrq.h MSG_APP)
        if not m["reject"]:
            li, lt = m["index"], m["logterm"]
            t = self.raftLog.terms
            if li < len(t):
                del t[li:]
            if li > len(t):
                t.extend([lt] * (li - len(t)))
            elif li > 0 and t[li - 1] != lt:
                t[li - 1] = lt
            self._commitTo(m["commit"])
        self._replyAck(m["from"])

    def handleHeartbeat(self, m):
        self._commitTo(m["commit"])
        self._replyAck(m["from"])

    def Step(self, m):
        ty = m["type"]
        if ty == MsgHup:
            if self.state != Leader:
                self.campaign()
            return
        if m["term"] == 0:
            pass
        elif m["term"] > self.Term:
            self.becomeFollower(m["term"], None_ if ty == MsgVote else m["from"])
        elif m["term"] < self.Term:
            return
        {Leader: self.stepLeader, Candidate: self.stepCandidate, Follower: self.stepFollower}[self.state](m)

    def stepLeader(self, m):
        ty = m["type"]
        if ty == MsgBeat:
            self.out |= OUT_BCAST_HEARTBEAT
        elif ty == MsgProp:
            self.appendEntry(m["n"])
            self.out |= OUT_BCAST_APPEND
        elif ty == MsgVote:
            self._replyVote(m["from"], True)
        elif ty == MsgAppResp:
            if not m["reject"] and self.prs[m["from"]].maybeUpdate(m["index"]):
                if self.maybeCommit():
                    self.out |= OUT_BCAST_APPEND

    def stepCandidate(self, m):
        ty = m["type"]
        if ty == MsgProp:
            self.out |= OUT_PROP_DROPPED
        elif ty == MsgApp:
            self.becomeFollower(self.Term, m["from"])
            self.handleAppendEntries(m)
        elif ty == MsgHeartbeat:
            self.becomeFollower(self.Term, m["from"])
            self.handleHeartbeat(m)
        elif ty == MsgVote:
            self._replyVote(m["from"], True)
        elif ty == MsgVoteResp:
            gr = self.poll(m["from"], not m["reject"])
            if self.q() == gr:
                self.becomeLeader()
                self.out |= OUT_BCAST_APPEND
            elif self.q() == len(self.votes) - gr:
                self.becomeFollower(self.Term, None_)

    def stepFollower(self, m):
        ty = m["type"]
        if ty == MsgProp:
            self.out |= OUT_PROP_DROPPED if self.lead == None_ else OUT_PROP_FORWARD
        elif ty == MsgApp:
            self.electionElapsed, self.lead = 0, m["from"]
            self.handleAppendEntries(m)
        elif ty == MsgHeartbeat:
            self.electionElapsed, self.lead = 0, m["from"]
            self.handleHeartbeat(m)
        elif ty == MsgVote:
            if (self.Vote == None_ or self.Vote == m["from"]) and self.raftLog.isUpToDate(m["index"], m["logterm"]):
                self.electionElapsed, self.Vote = 0, m["from"]
                self._replyVote(m["from"], False)
            else:
                self._replyVote(m["from"], True)

    def tick(self):
        if self.state == Leader:
            self.heartbeatElapsed += 1
            self.electionElapsed += 1
            if self.electionElapsed >= self.electionTimeout:
                self.electionElapsed = 0
            if self.heartbeatElapsed >= self.heartbeatTimeout:
                self.heartbeatElapsed = 0
                self.Step({"type": MsgBeat, "from": self.id, "term": 0})
        else:
            self.electionElapsed += 1
            if self.electionElapsed >= self.randomizedElectionTimeout:
                self.electionElapsed = 0
                self.Step({"type": MsgHup, "from": self.id, "term": 0})


class PyEngine:
    """G groups stepped per tick in the canonical order (senders ascending, proposals, Tick)."""

    def __init__(self, G, R, *, group_base=0, election_tick=10, heartbeat_tick=1, seed=0, self_id=0):
        self.G, self.R, self.tick_no = G, R, 0
        self.groups = [Raft(self_id or ((group_base + g) % R) + 1, range(1, R + 1), election_tick, heartbeat_tick, seed,
                            group_base + g, lambda: self.tick_no) for g in range(G)]

    def tick(self, ib):
        for g, r in enumerate(self.groups):
            r.out = 0
            for s in range(self.R):
                t = int(ib["type"][s, g])
                if (t & 0x0F) == 0 or s + 1 == r.id:
                    continue
                r.Step({"type": t & 0x0F, "reject": bool(t & 0x80), "from": s + 1, "term": int(ib["term"][s, g]),
                        "index": int(ib["index"][s, g]), "logterm": int(ib["logterm"][s, g]), "commit": int(ib["commit"][s, g])})
            n = int(ib["prop_count"][g])
            if n:
                r.Step({"type": MsgProp, "from": r.id, "term": 0, "n": n})
            r.tick()
        self.tick_no += 1

    def column(self, name):
        f = {"term": lambda r: r.Term, "vote": lambda r: r.Vote, "committed": lambda r: r.raftLog.committed,
             "last_index": lambda r: r.raftLog.lastIndex(), "last_term": lambda r: r.raftLog.lastTerm(),
             "role": lambda r: r.state, "lead": lambda r: r.lead, "election_elapsed": lambda r: r.electionElapsed,
             "heartbeat_elapsed": lambda r: r.heartbeatElapsed, "randomized_timeout": lambda r: r.randomizedElectionTimeout,
             "out": lambda r: r.out}[name]
        return [f(r) for r in self.groups]
