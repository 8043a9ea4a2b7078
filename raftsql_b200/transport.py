"""HTTP peer transport — the role etcd's rafthttp plays in the reference (reference raft.go:170-184 set-up,
:230 Send, :248-266 serveRaft, :268-270 Process): every node listens on its own peer URL (`peers[id-1]`,
raft.go:249) and POSTs batches of raft messages to the others.  SURVEY §8f row f4.

Same interface as hostnode.LocalTransport (register / unregister / send / drain), so a HostNode does not know
which one it has.  Unreachable peers lose messages — rafthttp reports them through ReportUnreachable, which the
reference implements as a no-op (raft.go:271-273).  The wire format is JSON (the reference's is protobuf
raftpb.Message over rafthttp streams; byte compatibility with etcd peers is out of scope — DESIGN.md §8).
"""
from __future__ import annotations

import json
import queue
import threading
import urllib.request
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from urllib.parse import urlparse

from .hostnode import Message


def encode(msgs) -> bytes:
    return json.dumps([[m.type, m.to, m.frm, m.term, m.logterm, m.index, m.commit, int(m.reject), m.reject_hint,
                        [[t, d.hex()] for (t, d) in m.entries]] for m in msgs], separators=(",", ":")).encode()


def decode(body: bytes):
    out = []
    for ty, to, frm, term, logterm, index, commit, rej, hint, ents in json.loads(body):
        out.append(Message(ty, to, frm, term=term, logterm=logterm, index=index, commit=commit, reject=bool(rej),
                           reject_hint=hint, entries=[(t, bytes.fromhex(d)) for t, d in ents]))
    return out


class HttpTransport:
    """One per process (or per node in the in-process tests)."""

    def __init__(self, peers):
        self.peers = list(peers)  # peers[i] is the URL of node id i+1 (raft.go:148-151,181-182)
        self.lock = threading.Lock()
        self.boxes: dict[int, list] = {}
        self.servers: dict[int, ThreadingHTTPServer] = {}
        self.senders: dict[int, queue.Queue] = {}
        self.closing = False

    # -- receiving: serveRaft (raft.go:248-266) + Process (raft.go:268-270) --------------------------------------
    def register(self, nid: int):
        u = urlparse(self.peers[nid - 1])
        tr = self

        class Handler(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, *a):
                pass

            def do_POST(self):
                n = int(self.headers.get("Content-Length") or 0)
                try:
                    msgs = decode(self.rfile.read(n))
                    with tr.lock:
                        box = tr.boxes.get(nid)
                        if box is not None:
                            box.extend(m for m in msgs if m.to == nid)
                    self.send_response(204)
                except Exception:
                    self.send_response(400)
                self.send_header("Content-Length", "0")
                self.end_headers()

        with self.lock:
            self.boxes[nid] = []
        srv = ThreadingHTTPServer((u.hostname, u.port), Handler)  # newStoppableListener (listener.go:25-59)
        srv.daemon_threads = True
        self.servers[nid] = srv
        threading.Thread(target=srv.serve_forever, daemon=True, name=f"serveRaft-{nid}").start()

    def unregister(self, nid: int):
        with self.lock:
            self.boxes.pop(nid, None)
        srv = self.servers.pop(nid, None)
        if srv is not None:  # stopHTTP (raft.go:198-202)
            srv.shutdown()
            srv.server_close()

    def drain(self, nid: int):
        with self.lock:
            box = self.boxes.get(nid)
            if box is None:
                return []
            out, box[:] = list(box), []
            return out

    # -- sending: transport.Send(rd.Messages) (raft.go:230) -----------------------------------------------------------
    def send(self, msgs):
        by_dest: dict[int, list] = {}
        for m in msgs:
            by_dest.setdefault(m.to, []).append(m)
        for to, batch in by_dest.items():
            with self.lock:
                local = self.boxes.get(to)
                if local is not None and to in self.servers:  # same process: short-circuit
                    local.extend(batch)
                    continue
            try:
                self._sender(to).put_nowait(batch)
            except queue.Full:
                pass  # a peer that has been unreachable for a while: drop, never stall the node's own tick loop

    def _sender(self, to: int) -> queue.Queue:
        with self.lock:
            q = self.senders.get(to)
            if q is None:
                q = self.senders[to] = queue.Queue(maxsize=256)
                threading.Thread(target=self._pump, args=(to, q), daemon=True, name=f"peer-{to}").start()
            return q

    def _pump(self, to: int, q: queue.Queue):
        url = self.peers[to - 1].rstrip("/") + "/raft"
        while not self.closing:
            try:
                batch = q.get(timeout=0.2)
            except queue.Empty:
                continue
            while True:  # coalesce whatever queued up meanwhile into one POST
                try:
                    batch = batch + q.get_nowait()
                except queue.Empty:
                    break
            try:
                req = urllib.request.Request(url, data=encode(batch), method="POST")
                urllib.request.urlopen(req, timeout=1.0).read()
            except Exception:
                pass  # unreachable peer: the messages are lost; raft retries (ReportUnreachable is a no-op upstream too)

    def close(self):
        self.closing = True
        for nid in list(self.servers):
            self.unregister(nid)
