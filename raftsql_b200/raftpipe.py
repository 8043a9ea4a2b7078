"""raftpipe — the drop-in seam, kept in the reference's shape (reference raftpipe.go:3-17).

    rp = NewRaftPipe(id, peers, proposeC)      # raftpipe.go:9-12
    rp.ProposeC  (send SQL strings)            rp.CommitC  (receive committed strings, None = "log replayed")
    rp.ErrorC                                   rp.Close()  # raftpipe.go:14-17: close(ProposeC); return <-ErrorC

Channel protocol preserved from the reference: all three channels unbuffered (raft.go:65-66);
on start every replayed entry, then one None, then live entries (raft.go:57-61,130-132); entries in log order,
empty/conf-change entries skipped (raft.go:84-86); shutdown = caller closes ProposeC -> CommitC closed, ErrorC
closed with no value so Close() returns None (raft.go:216-217,241-243,191-196).

Below the seam, `newRaftNode` (raft.go:62-78) builds a HostNode around a GPU engine of G = 1 group and
R = len(peers) replicas and runs the `serveChannels` loop (raft.go:204-246) on one thread.
"""
from __future__ import annotations

import threading
import time

from .hostnode import HostNode, LocalTransport


class ChanClosed(Exception):
    pass


class Chan:
    """A Go channel: unbuffered rendezvous by default, close(), `v, ok = recv()`, iteration until closed."""

    def __init__(self, buffered: int = 0):
        self.cap = buffered
        self.cv = threading.Condition()
        self.items: list = []  # [(ticket, value)]: a handful at most, one per blocked sender
        self.closed = False
        self.seq = 0

    def send(self, v, stop: threading.Event | None = None) -> bool:
        """Blocks until a receiver has taken the value (unbuffered) or there is room (buffered).  With `stop` it is
        `select { case ch <- v: case <-stopc: }` (raft.go:89-93): an aborted send WITHDRAWS its value — as in Go the
        receiver never sees it — and returns False.  A channel closed under a blocked sender raises (Go panics)."""
        with self.cv:
            if self.closed:
                raise ChanClosed("send on closed channel")
            self.seq += 1
            my = self.seq
            self.items.append((my, v))
            self.cv.notify_all()
            if self.cap and len(self.items) <= self.cap:
                return True
            while any(t == my for t, _ in self.items):  # unbuffered: until a receiver has taken this value
                if self.closed:
                    self._withdraw(my)
                    raise ChanClosed("channel closed while sending")
                if stop is not None and stop.is_set():
                    self._withdraw(my)
                    return False
                self.cv.wait(0.02 if stop is not None else 0.05)
            return True

    def _withdraw(self, ticket):
        self.items = [(t, v) for t, v in self.items if t != ticket]

    def recv(self, timeout: float | None = None):
        """-> (value, ok); ok is False when the channel is closed and drained."""
        end = None if timeout is None else time.monotonic() + timeout
        with self.cv:
            while not self.items:
                if self.closed:
                    return None, False
                if end is not None:
                    left = end - time.monotonic()
                    if left <= 0:
                        raise TimeoutError("recv timed out")
                    self.cv.wait(min(left, 0.05))
                else:
                    self.cv.wait(0.05)
            _, v = self.items.pop(0)
            self.cv.notify_all()
            return v, True

    def try_recv(self):
        """`select { case v, ok := <-ch: ... default: }` -> (value, ok, ready): ready is False when nothing is
        offered right now; ok is False (with ready True) once the channel is closed and drained."""
        with self.cv:
            if self.items:
                _, v = self.items.pop(0)
                self.cv.notify_all()
                return v, True, True
            return (None, False, True) if self.closed else (None, False, False)

    def close(self):
        with self.cv:
            self.closed = True
            self.cv.notify_all()

    def __iter__(self):
        while True:
            v, ok = self.recv()
            if not ok:
                return
            yield v


_registry_lock = threading.Lock()
_transports: dict[tuple, LocalTransport] = {}


def transport_for(peers) -> LocalTransport:
    """Nodes of one process that were given the same peer list share an in-process transport — the role
    loopback TCP + rafthttp play in the reference's in-process test cluster (raftsql_test.go:16-41)."""
    key = tuple(peers)
    with _registry_lock:
        if key not in _transports:
            _transports[key] = LocalTransport()
        return _transports[key]


class RaftPipe:
    def __init__(self, ProposeC: Chan, CommitC: Chan, ErrorC: Chan, node_thread=None):
        self.ProposeC, self.CommitC, self.ErrorC = ProposeC, CommitC, ErrorC
        self._thread = node_thread

    def Close(self):
        """reference raftpipe.go:14-17"""
        self.ProposeC.close()
        err, ok = self.ErrorC.recv()
        if self._thread is not None:
            self._thread.join(timeout=10)
        return err if ok else None


def make_engine_core(npeers: int, nid: int, *, device: int = 0, seed: int = 0, election_tick: int = 10,
                     heartbeat_tick: int = 1):
    """The product core: one GPU engine holding this node's replica of the (single) group."""
    from .engine import Engine

    return Engine(1, npeers, self_id=nid, device=device, seed=seed or (0x5EED + nid), election_tick=election_tick,
                  heartbeat_tick=heartbeat_tick, inbox_slots=1)


def newRaftNode(id: int, peers, proposeC: Chan, *, tick_seconds: float = 0.1, waldir: str | None = "auto",
                core_factory=make_engine_core, transport: LocalTransport | None = None, **core_kw):
    """reference raft.go:62-78.  Returns (commitC, errorC, thread).  tick_seconds defaults to the reference's
    100 ms ticker (raft.go:207); ElectionTick 10 / HeartbeatTick 1 (raft.go:154-155) are the engine defaults."""
    commitC, errorC = Chan(), Chan()
    tr = transport or transport_for(peers)
    core = core_factory(len(peers), id, **core_kw)
    wd = f"raftsql-{id}" if waldir == "auto" else waldir  # raft.go:69
    node = HostNode(core, id, len(peers), tr, wd)
    stop = threading.Event()

    def feed():  # raft.go:211-218: proposals -> raft; closing proposeC shuts the node down
        for prop in proposeC:
            with plock:
                node.propose(prop.encode() if isinstance(prop, str) else bytes(prop))
        stop.set()

    plock = threading.Lock()

    def publish(payloads) -> bool:  # raft.go:82-96
        for d in payloads:
            s = d.decode()
            while True:
                if stop.is_set():
                    return False
                try:
                    _send_or_stop(commitC, s, stop)
                    break
                except ChanClosed:
                    return False
        return True

    def run():  # startRaft (raft.go:144-188) + serveChannels (raft.go:204-246)
        err = None
        try:
            node.start()
            if not publish(node.replay):
                return
            _send_or_stop(commitC, None, stop)  # "commit channel is current" (raft.go:131-132)
            feeder = threading.Thread(target=feed, daemon=True)
            feeder.start()
            nxt = time.monotonic()
            while not stop.is_set():
                with plock:
                    out = node.step_tick()
                if not publish(out):
                    break
                nxt += tick_seconds
                delay = nxt - time.monotonic()
                if delay > 0:
                    stop.wait(delay)
                else:
                    nxt = time.monotonic()
        except ChanClosed:
            pass
        except Exception as ex:  # writeError (raft.go:136-142): commitC closed, then the error, then errorC closed
            err = ex
        finally:
            node.stop()
            try:
                core.close()
            except Exception:
                pass
            commitC.close()
            if err is not None:
                try:
                    errorC.send(err)
                except ChanClosed:
                    pass
            errorC.close()

    th = threading.Thread(target=run, daemon=True, name=f"raftnode-{id}")
    th.node = node  # introspection for tests / operators (role, term, commit): not part of the seam
    th.start()
    return commitC, errorC, th


def _send_or_stop(ch: Chan, v, stop: threading.Event):
    """`select { case commitC <- v: case <-stopc: }` (raft.go:89-93); raises ChanClosed when the stop side won."""
    if not ch.send(v, stop):
        raise ChanClosed()


def NewRaftPipe(id: int, peers, proposeC: Chan, **kw) -> RaftPipe:
    """reference raftpipe.go:9-12"""
    cC, eC, th = newRaftNode(id, peers, proposeC, **kw)
    return RaftPipe(proposeC, cC, eC, th)
