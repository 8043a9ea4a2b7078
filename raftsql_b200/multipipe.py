"""multipipe — many raft groups of ONE node behind per-group raftPipe seams over ONE engine (SURVEY §8b "Go shim …
a NewMultiRaftPipe(groups …) variant demuxes committed[g] advances into per-group CommitCs", §8f row f1).

    mp = NewMultiRaftPipe(id, peers, n_groups)          # one engine of G = n_groups groups, R = len(peers)
    mp.ProposeC[g].send("INSERT ...")                   # per-group seams, each with the protocol of raftpipe.go:3-17:
    mp.CommitC[g]   -> replayed entries, then None, then live entries of group g, in log order
    mp.ErrorC       -> one for the node;  mp.Close() closes every ProposeC and returns <-ErrorC

The reference runs one raft group per process (raft.go:62-78); this is the shape the engine was built for — the
per-tick arithmetic of all G groups is ONE `mrq_tick` (one kernel pass), not G of them.  Every group keeps the
validated single-group host logic (`hostnode.HostNode`: log, maybeAppend, Progress.Next, Ready.Messages, WAL), wired
to a per-group VIEW of the shared engine: the views collect every group's inbox messages and proposals, the node
posts them in one `mrq_post_inbox_delta` + one `mrq_propose`, ticks once, exports the state columns once, and each
group finishes its Ready handling from its own slice.  Durability is batched the same way: the groups append to ONE
write-ahead file (`MultiWal`), the node fsyncs it ONCE per tick, and only then do that tick's messages leave
(persist before send, raft.go:228-230) — a group-commit WAL instead of one fsync per writing group.
"""
from __future__ import annotations

import json
import os
import struct
import threading
import time

import numpy as np

from . import hostnode
from .hostnode import HostNode
from .raftpipe import Chan, ChanClosed, _send_or_stop

STATE_COLS = ("term", "vote", "committed", "last_index", "last_term", "term_start", "match", "role", "lead")


class MultiLocalTransport:
    """In-process stand-in for rafthttp with one mailbox per (node, group); `group(g)` is the LocalTransport
    interface a HostNode expects, scoped to group g."""

    def __init__(self):
        self.lock = threading.Lock()
        self.boxes: dict[tuple[int, int], list] = {}

    def group(self, g: int):
        return _GroupTransport(self, g)


class _GroupTransport:
    def __init__(self, tr: MultiLocalTransport, g: int):
        self.tr, self.g = tr, g

    def register(self, nid: int):
        with self.tr.lock:
            self.tr.boxes[(nid, self.g)] = []

    def unregister(self, nid: int):
        with self.tr.lock:
            self.tr.boxes.pop((nid, self.g), None)

    def send(self, msgs):
        with self.tr.lock:
            for m in msgs:
                box = self.tr.boxes.get((m.to, self.g))
                if box is not None:  # unknown / stopped peers lose messages, like a dead TCP peer
                    box.append(m)

    def drain(self, nid: int):
        with self.tr.lock:
            box = self.tr.boxes.get((nid, self.g))
            if box is None:
                return []
            out, box[:] = list(box), []
            return out


class MultiWal:
    """Group-commit write-ahead log: ONE append-only file for all groups of a node (`<dir>/wal.log`, the records of
    hostnode.Wal tagged with their group), made durable with ONE fsync per tick however many groups wrote — a
    per-group WAL would cost one fsync per writing group per tick.  `view(g)` is the hostnode.Wal interface scoped
    to group g; its `save()` only appends, `MultiHostNode` calls `sync()` once per tick before anything is sent."""

    def __init__(self, dirname: str):
        self.dir = dirname
        self.path = os.path.join(dirname, "wal.log")
        self.f = None
        self.dirty = False
        self.syncs = 0
        self._parsed = None

    def open(self):
        if self.f is None:
            os.makedirs(self.dir, mode=0o750, exist_ok=True)
            self.parse()                      # replay BEFORE the tail is cut and new records land
            hostnode.repair_tail(self.path)   # a torn tail is cut off, or every later record would hide behind it
            self.f = open(self.path, "ab")

    def parse(self) -> dict:
        """-> {g: (hardstate or None, [(term, data)])}, tolerant of a torn tail like hostnode.Wal.read_all"""
        if self._parsed is not None:
            return self._parsed
        out: dict = {}
        for rec in hostnode.scan_records(self.path)[0]:
            hs, ents = out.setdefault(rec["g"], [None, []])
            if "hs" in rec:
                out[rec["g"]][0] = tuple(rec["hs"])
            elif "e" in rec:
                i, t, d = rec["e"]
                del ents[i - 1:]
                ents.append((t, bytes.fromhex(d)))
            elif "t" in rec:
                del ents[rec["t"]:]
        self._parsed = {g: (v[0], v[1]) for g, v in out.items()}
        return self._parsed

    def put(self, rec: dict):
        self.f.write(hostnode.frame_record(rec))
        self.dirty = True

    def sync(self):
        """the one fsync of the tick (wal.Save's durability point, raft.go:228, for every group at once)"""
        if self.f is not None and self.dirty:
            self.f.flush()
            os.fsync(self.f.fileno())
            self.dirty = False
            self.syncs += 1

    def close(self):
        if self.f is not None:
            self.sync()
            self.f.close()
            self.f = None

    def view(self, g: int):
        return _GroupWal(self, g)


class _GroupWal:
    """hostnode.Wal's interface for one group of a MultiWal"""

    def __init__(self, wal: MultiWal, g: int):
        self.wal, self.g, self.dir = wal, g, wal.dir

    def open(self):
        self.wal.open()

    def read_all(self):
        hs, ents = self.wal.parse().get(self.g, (None, []))
        return hs, list(ents)

    def save(self, hardstate, new_entries, first_index, truncate_after=None):
        if truncate_after is not None:
            self.wal.put({"g": self.g, "t": truncate_after})
        for k, (t, d) in enumerate(new_entries):
            self.wal.put({"g": self.g, "e": [first_index + k, t, d.hex()]})
        if hardstate is not None:
            self.wal.put({"g": self.g, "hs": list(hardstate)})

    def close(self):
        pass  # the owner closes the shared file


class _DeferredTransport:
    """A group's transport whose sends wait in the node's outbox until the tick's WAL records are durable
    (persist before send: wal.Save precedes transport.Send, raft.go:228-230)."""

    def __init__(self, tr, owner: "MultiHostNode"):
        self.tr, self.owner = tr, owner

    def register(self, nid):
        self.tr.register(nid)

    def unregister(self, nid):
        self.tr.unregister(nid)

    def drain(self, nid):
        return self.tr.drain(nid)

    def send(self, msgs):
        self.owner.outbox.append((self.tr, list(msgs)))


class _GroupCore:
    """What HostNode needs from its consensus core, for ONE group of the shared engine: posts are collected by the
    owning MultiHostNode, reads come from the columns it exported after the shared tick."""

    def __init__(self, owner: "MultiHostNode", g: int):
        self.owner, self.g = owner, g
        self.hardstate = None

    def import_state(self, st: dict):  # HostNode.start(): restored HardState + log position of this group
        self.hardstate = {k: int(np.asarray(v).reshape(-1)[0]) for k, v in st.items()}

    def post_inbox_delta(self, msgs, slot=0, accumulate=False):
        self.owner.batch_msgs.extend((self.g,) + tuple(m[1:]) for m in msgs)

    def propose(self, groups, counts, slot=0):
        self.owner.batch_props.append((self.g, int(counts[0])))

    def tick(self, slot=0):
        raise RuntimeError("a group view does not tick: MultiHostNode ticks the shared engine once for all groups")

    def export_state(self, columns=None):
        s, g = self.owner.state, self.g
        return {k: (v[:, g:g + 1] if v.ndim == 2 else v[g:g + 1]) for k, v in s.items()}

    def sync_out(self):
        return self.owner.out[self.g:self.g + 1]


class MultiHostNode:
    """One node's replicas of G groups: G HostNodes around one engine core."""

    def __init__(self, core, nid: int, npeers: int, n_groups: int, transport: MultiLocalTransport, waldir: str | None = None):
        self.core, self.id, self.G = core, nid, n_groups
        self.batch_msgs: list = []
        self.batch_props: list = []
        self.state: dict = {}
        self.out = np.zeros(n_groups, np.uint32)
        self.views = [_GroupCore(self, g) for g in range(n_groups)]
        self.outbox: list = []
        self.wal = MultiWal(waldir) if waldir else None  # group commit: one file, one fsync per tick
        self.nodes = [HostNode(self.views[g], nid, npeers, _DeferredTransport(transport.group(g), self),
                               self.wal.view(g) if self.wal else None) for g in range(n_groups)]

    def start(self) -> list[list[bytes]]:
        """replayWAL for every group (raft.go:122-134), then ONE import of the restored columns into the engine.
        Returns the committed payloads to replay, per group."""
        cols = {k: np.zeros(self.G, np.uint64) for k in ("term", "vote", "committed", "last_index", "last_term")}
        for g, n in enumerate(self.nodes):
            n.start()
            hs = self.views[g].hardstate
            if hs:
                for k in cols:
                    cols[k][g] = hs.get(k, 0)
        if any(v.hardstate for v in self.views):
            self.core.import_state(cols)
        return [n.replay for n in self.nodes]

    def propose(self, g: int, data: bytes):
        self.nodes[g].propose(data)

    def step_tick(self) -> list[list[bytes]]:
        """One tick of every group: per-group prepare, ONE engine tick, per-group Ready.  Returns, per group, the
        payloads newly committed (what goes to that group's CommitC)."""
        self.batch_msgs, self.batch_props = [], []
        replies = [n.prepare_tick() for n in self.nodes]
        self.core.post_inbox_delta(self.batch_msgs, slot=0)
        if self.batch_props:
            self.core.propose([g for g, _ in self.batch_props], [c for _, c in self.batch_props], slot=0)
        self.core.tick(0)
        self.state = self.core.export_state(STATE_COLS)
        self.out = self.core.sync_out()
        published = [n.finish_tick(r) for n, r in zip(self.nodes, replies)]
        self.flush()
        return published

    def flush(self):
        """make this tick's WAL records of every group durable with one fsync, THEN let the messages out"""
        if self.wal is not None:
            self.wal.sync()
        outbox, self.outbox = self.outbox, []
        for tr, msgs in outbox:
            tr.send(msgs)

    def stop(self):
        for n in self.nodes:
            n.stop()
        if self.wal is not None:
            self.wal.close()


class MultiRaftPipe:
    def __init__(self, ProposeC, CommitC, ErrorC: Chan, thread=None):
        self.ProposeC, self.CommitC, self.ErrorC = ProposeC, CommitC, ErrorC
        self._thread = thread

    def Close(self):
        """raftpipe.go:14-17, for every group: close(ProposeC[g]); return <-ErrorC"""
        for c in self.ProposeC:
            c.close()
        err, ok = self.ErrorC.recv()
        if self._thread is not None:
            self._thread.join(timeout=10)
        return err if ok else None


def make_engine_core(npeers: int, nid: int, n_groups: int, *, device: int = 0, seed: int = 0, election_tick: int = 10,
                     heartbeat_tick: int = 1):
    """The product core: one GPU engine holding this node's replica of every group."""
    from .engine import Engine

    return Engine(n_groups, npeers, self_id=nid, device=device, seed=seed or (0x5EED + nid), election_tick=election_tick,
                  heartbeat_tick=heartbeat_tick, inbox_slots=1)


def NewMultiRaftPipe(id: int, peers, n_groups: int, proposeCs=None, *, tick_seconds: float = 0.1, waldir: str | None = "auto",
                     core_factory=make_engine_core, transport: MultiLocalTransport | None = None, commit_buffer: int = 0,
                     **core_kw) -> MultiRaftPipe:
    """One node of a multi-raft cluster.  `proposeCs`: one Chan per group (created if None).  Channels are
    unbuffered like the reference's (raft.go:65-66); `commit_buffer` > 0 buffers the CommitCs for hosts that drain
    many groups from few threads."""
    proposeCs = list(proposeCs) if proposeCs is not None else [Chan() for _ in range(n_groups)]
    assert len(proposeCs) == n_groups
    commitCs = [Chan(commit_buffer) for _ in range(n_groups)]
    errorC = Chan()
    tr = transport or _shared_transport(tuple(peers), n_groups)
    core = core_factory(len(peers), id, n_groups, **core_kw)
    wd = f"raftsql-{id}" if waldir == "auto" else waldir  # raft.go:69; one group-commit file for all groups (MultiWal)
    node = MultiHostNode(core, id, len(peers), n_groups, tr, wd)
    stop = threading.Event()

    def publish(g, payloads) -> bool:  # raft.go:82-96
        for d in payloads:
            if stop.is_set():
                return False
            try:
                _send_or_stop(commitCs[g], d.decode(), stop)
            except ChanClosed:
                return False
        return True

    def feed() -> bool:
        """raft.go:211-218 for every group, without a goroutine per group: take whatever is offered right now;
        the node shuts down once every ProposeC has been closed."""
        open_any = False
        for g, pc in enumerate(proposeCs):
            while True:
                v, ok, ready = pc.try_recv()
                if not ready:
                    open_any = True
                    break
                if not ok:
                    break  # closed and drained
                node.propose(g, v.encode() if isinstance(v, str) else bytes(v))
        return open_any

    def run():  # startRaft (raft.go:144-188) + serveChannels (raft.go:204-246), all groups on one thread
        err = None
        try:
            for g, replay in enumerate(node.start()):
                if not publish(g, replay):
                    return
                _send_or_stop(commitCs[g], None, stop)  # "commit channel is current" (raft.go:131-132)
            nxt = time.monotonic()
            while not stop.is_set():
                if not feed():
                    stop.set()
                    break
                for g, out in enumerate(node.step_tick()):
                    if not publish(g, out):
                        return
                nxt += tick_seconds
                delay = nxt - time.monotonic()
                if delay > 0:
                    stop.wait(delay)
                else:
                    nxt = time.monotonic()
        except ChanClosed:
            pass
        except Exception as ex:  # writeError (raft.go:136-142): commitCs closed, then the error, then errorC closed
            err = ex
        finally:
            stop.set()
            node.stop()
            try:
                core.close()
            except Exception:
                pass
            for c in commitCs:
                c.close()
            if err is not None:
                try:
                    errorC.send(err)
                except ChanClosed:
                    pass
            errorC.close()

    th = threading.Thread(target=run, daemon=True, name=f"multiraft-{id}")
    th.node = node  # introspection for tests / operators: not part of the seam
    th.start()
    return MultiRaftPipe(proposeCs, commitCs, errorC, th)


_registry_lock = threading.Lock()
_transports: dict[tuple, MultiLocalTransport] = {}


def _shared_transport(key_peers: tuple, n_groups: int) -> MultiLocalTransport:
    """Nodes of one process that were given the same peer list share an in-process transport."""
    with _registry_lock:
        return _transports.setdefault((key_peers, n_groups), MultiLocalTransport())
