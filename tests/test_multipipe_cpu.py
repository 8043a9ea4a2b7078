"""The multi-group seam (raftsql_b200.multipipe): G raft groups of one node behind per-group
ProposeC / CommitC pairs over ONE engine, ticked once per tick for all groups.  The reference's two tests
(raftsql_test.go:92-171) restated per group on a 3-node in-process cluster: every group elects a leader, entries
proposed through any node commit on every node in that group's log order, groups do not leak into each other, the
nil sentinel arrives once per group, and a stopped node replays every group's WAL and catches up.

CPU: the oracle as the G-group core (tests/oracle_core.py).  The same host code over the GPU engine is one
`core_factory` away (multipipe.make_engine_core)."""
import threading
import time

from oracle_core import make_oracle_multicore
from raftsql_b200.multipipe import MultiLocalTransport, NewMultiRaftPipe

PEERS = ["http://127.0.0.1:10000", "http://127.0.0.1:10001", "http://127.0.0.1:10002"]


class Collector:
    """drains one CommitC on its own thread, like db.go's readCommits"""

    def __init__(self, ch):
        self.ch, self.got, self.nils, self.lock = ch, [], 0, threading.Lock()
        self.th = threading.Thread(target=self.run, daemon=True)
        self.th.start()

    def run(self):
        for v in self.ch:
            with self.lock:
                if v is None:
                    self.nils += 1
                else:
                    self.got.append(v)

    def snapshot(self):
        with self.lock:
            return list(self.got)


def wait_until(pred, timeout=40.0):
    end = time.monotonic() + timeout
    while time.monotonic() < end:
        if pred():
            return True
        time.sleep(0.01)
    return False


def uniq(v):
    out = []
    for x in v:
        if x not in out:
            out.append(x)
    return out


class Cluster:
    def __init__(self, G, tmp, tick=0.005):
        self.G, self.tmp, self.tick = G, str(tmp), tick
        self.tr = MultiLocalTransport()
        self.mp, self.col = [None] * 3, [None] * 3
        for i in range(3):
            self.start(i)

    def start(self, i):
        self.mp[i] = NewMultiRaftPipe(i + 1, PEERS, self.G, tick_seconds=self.tick, core_factory=make_oracle_multicore,
                                      transport=self.tr, waldir=f"{self.tmp}/raftsql-{i + 1}")
        self.col[i] = [Collector(c) for c in self.mp[i].CommitC]

    def stop(self, i):
        assert self.mp[i].Close() is None
        self.mp[i] = None

    def close(self):
        for i in range(3):
            if self.mp[i] is not None:
                self.stop(i)

    def alive(self):
        return [i for i in range(3) if self.mp[i] is not None]

    def leader_of(self, g):
        for i in self.alive():
            role = self.mp[i]._thread.node.state.get("role")
            if role is not None and int(role[g]) == 2:
                return i
        return None

    def propose_until_committed(self, g, text, via):
        """A raft client retries: an entry accepted by a leader that is deposed before replicating it is lost
        (upstream too), so proposals are at-least-once and the applied sequence is read modulo repeats.  Retries
        only ever fire when the machine stalls for seconds; they keep the scenario meaningful instead of flaky."""
        for attempt in range(8):
            ld = self.leader_of(g)
            self.mp[via if attempt == 0 or ld is None else ld].ProposeC[g].send(text)
            if wait_until(lambda: all(text in self.col[i][g].snapshot() for i in self.alive()), timeout=5.0):
                return True
        return False


def test_every_group_replicates_independently_and_in_order(tmp_path):
    G = 6
    c = Cluster(G, tmp_path)
    try:
        # group g gets its own sequence, first through node g % 3 (forwarded to whoever leads); groups interleave
        want = {g: [f"g{g}-entry-{k}" for k in range(5)] for g in range(G)}
        for k in range(5):
            for g in range(G):
                assert c.propose_until_committed(g, want[g][k], g % 3), f"group {g} entry {k}"
        for g in range(G):
            for i in range(3):
                assert uniq(c.col[i][g].snapshot()) == want[g], f"node {i} group {g}"  # log order, nothing from other groups
                assert c.col[i][g].snapshot() == c.col[0][g].snapshot()  # every node applied the same sequence
                assert c.col[i][g].nils == 1  # "commit channel is current", once per group
        # one engine tick served all groups: every group has exactly one leader somewhere in the cluster
        assert wait_until(lambda: all(sum(int(c.mp[i]._thread.node.state["role"][g]) == 2 for i in range(3)) == 1 for g in range(G)))
    finally:
        c.close()


def test_stopped_node_replays_every_groups_wal_and_catches_up(tmp_path):
    G = 4
    c = Cluster(G, tmp_path)
    try:
        for g in range(G):
            assert c.propose_until_committed(g, f"CREATE-{g}", 0)
        # stop node 2 (index 1): the groups it led re-elect among the remaining two and keep committing with 2 of 3
        c.stop(1)
        for g in range(G):
            assert wait_until(lambda: c.leader_of(g) is not None), f"group {g} must re-elect with 2 of 3 nodes"
            assert c.propose_until_committed(g, f"while-down-{g}", c.leader_of(g) or 0), f"group {g} must commit with 2 of 3 nodes"
        c.start(1)
        for g in range(G):
            assert wait_until(lambda: f"while-down-{g}" in c.col[1][g].snapshot()), f"group {g} on the restarted node: {c.col[1][g].snapshot()}"
            assert uniq(c.col[1][g].snapshot()) == [f"CREATE-{g}", f"while-down-{g}"]  # replay, then catch-up, in order
            assert c.col[1][g].nils == 1
    finally:
        c.close()


def test_group_commit_wal_one_fsync_per_tick_and_persist_before_send(tmp_path, monkeypatch):
    """The node's WAL is ONE file for all groups, made durable once per tick (a per-group WAL would fsync once per
    writing group), and nothing leaves the node before that fsync (wal.Save precedes transport.Send, raft.go:228-230)."""
    import os

    from raftsql_b200 import multipipe
    from raftsql_b200.hostnode import Message
    from raftsql_b200 import _ffi as F

    G = 5
    fsyncs = []
    real_fsync = os.fsync
    monkeypatch.setattr(os, "fsync", lambda fd: (fsyncs.append(fd), real_fsync(fd))[1])
    tr = MultiLocalTransport()
    for g in range(G):  # peers 2 and 3 exist as mailboxes only
        tr.group(g).register(2), tr.group(g).register(3)
    core = make_oracle_multicore(3, 1, G)
    node = multipipe.MultiHostNode(core, 1, 3, G, tr, str(tmp_path / "raftsql-1"))
    sent_while_dirty = []
    real_send = multipipe._GroupTransport.send

    def checked_send(self, msgs):
        if msgs and node.wal.dirty:
            sent_while_dirty.append(msgs)
        return real_send(self, msgs)

    monkeypatch.setattr(multipipe._GroupTransport, "send", checked_send)
    node.start()
    ticks = 0
    for _ in range(60):  # every group times out and campaigns at its own tick: HardState changes -> WAL records
        before = len(fsyncs)
        node.step_tick()
        ticks += 1
        assert len(fsyncs) - before <= 1, "at most one fsync per tick, however many groups wrote"
        for g in range(G):  # node 2 grants every vote request it sees
            for m in tr.group(g).drain(2):
                if m.type == F.MSG_VOTE:
                    tr.group(g).send([Message(F.MSG_VOTE_RESP, 1, 2, term=m.term)])
            tr.group(g).drain(3)
    roles = node.state["role"]
    assert all(int(r) == 2 for r in roles), "every group elected this node with node 2's vote"
    assert not sent_while_dirty, "a message left the node before its tick's WAL records were durable"
    assert 0 < len(fsyncs) <= ticks and len(fsyncs) < G * 3, f"{len(fsyncs)} fsyncs for {G} groups over {ticks} ticks"
    node.stop()
    # and the one file replays every group: term, vote for self, the leader's empty entry
    again = multipipe.MultiWal(str(tmp_path / "raftsql-1")).parse()
    assert sorted(again) == list(range(G))
    for g in range(G):
        hs, ents = again[g]
        assert hs[0] >= 1 and hs[1] == 1 and len(ents) >= 1 and ents[0][1] == b""


def test_close_protocol(tmp_path):
    """Close() closes every ProposeC; every CommitC is closed; ErrorC closes with no value -> None"""
    c = Cluster(3, tmp_path)
    mp = c.mp[0]
    assert mp.Close() is None
    for ch in mp.CommitC:
        assert wait_until(lambda: ch.closed)
    c.mp[0] = None
    c.close()


def test_core_failure_reaches_errorc(tmp_path):
    """writeError (raft.go:136-142): CommitCs closed, then the error on ErrorC, then ErrorC closed"""

    class Boom(RuntimeError):
        pass

    def bad_core(npeers, nid, n_groups, **kw):
        core = make_oracle_multicore(npeers, nid, n_groups, **kw)
        real, n = core.tick, [0]

        def tick(slot=0):
            n[0] += 1
            if n[0] == 5:
                raise Boom("engine fault")
            return real(slot)

        core.tick = tick
        return core

    mp = NewMultiRaftPipe(1, PEERS[:1], 2, tick_seconds=0.005, core_factory=bad_core, transport=MultiLocalTransport(),
                          waldir=None)
    cols = [Collector(ch) for ch in mp.CommitC]
    err, ok = mp.ErrorC.recv(timeout=20)
    assert ok and isinstance(err, Boom)
    assert mp.ErrorC.recv(timeout=5) == (None, False)
    for col in cols:
        col.th.join(5)
        assert col.ch.closed
