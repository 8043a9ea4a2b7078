"""hostnode — the host side of one raft node: log, WAL, message construction, commit publishing.

This is what surrounds the hot path in the reference's `raftNode` (reference raft.go:38-273): the
`serveChannels` loop (raft.go:204-246) — tick, drain Ready, persist, send, publish — with the consensus
arithmetic (everything etcd-raft's `node.Tick/Step/Propose/Ready` computed, raft.go:214,224,227,269) done
by the engine core on the GPU.  SURVEY §8f rows f2 (plumbing), f3 (durable state) and f4 (inbox builder).

The core is anything with the Engine's tick protocol (`post_inbox_delta / propose / tick / export_state /
sync_out`); the product passes `raftsql_b200.Engine`.  The node keeps what the engine deliberately does not:
entry payloads and per-entry terms (the log), follower-side log matching (`maybeAppend`), `Progress.Next`
bookkeeping for `sendAppend`, and the WAL.

Divergences from the reference, all deliberate and recorded in DESIGN.md §8:
  * entries are published on commit (index <= committed), not on local append (raft.go:231 publishes
    rd.Entries — SURVEY §3.2 "reference quirk");
  * HardState is restored on restart (the reference discards it, raft.go:124);
  * membership is the static peer list at every start (the reference loses it on restart).
"""
from __future__ import annotations

import json
import os
import struct
import threading
import zlib
from dataclasses import dataclass, field

import numpy as np

from . import _ffi as F

MsgProp = 2  # raftpb.MsgProp: host-level only (forwarded proposals); never enters the engine inbox


@dataclass
class Message:  # raftpb.Message, the fields this system uses
    type: int
    to: int
    frm: int
    term: int = 0
    logterm: int = 0
    index: int = 0
    commit: int = 0
    reject: bool = False
    reject_hint: int = 0
    entries: list = field(default_factory=list)  # [(term, payload bytes)]


class Log:
    """raft.MemoryStorage + raftLog's entry half (reference raft.go:70,129,229): index-1 based, no compaction
    (the reference never compacts: SURVEY §5)."""

    def __init__(self):
        self.ents: list[tuple[int, bytes]] = []  # ents[i-1] = (term, data) of entry i

    def last_index(self) -> int:
        return len(self.ents)

    def term(self, i: int) -> int:
        return self.ents[i - 1][0] if 1 <= i <= len(self.ents) else 0

    def last_term(self) -> int:
        return self.term(len(self.ents))

    def slice(self, lo: int, hi: int | None = None, max_bytes: int = 1 << 20):
        """entries [lo, hi]; limited like raft.Config.MaxSizePerMsg (reference raft.go:157, 1 MiB)."""
        hi = len(self.ents) if hi is None else hi
        out, size = [], 0
        for i in range(lo, hi + 1):
            t, d = self.ents[i - 1]
            size += len(d)
            if out and size > max_bytes:
                break
            out.append((t, d))
        return out

    def maybe_append(self, index: int, logterm: int, committed: int, ents: list):
        """upstream raftLog.maybeAppend: (lastnewi, ok, first_written, truncated).  Conflicts truncate the
        suffix (findConflict); first_written is the index of the first entry this call wrote (0: none)."""
        if self.term(index) != logterm and not (index == 0 and logterm == 0):
            return 0, False, 0, False
        lastnewi = index + len(ents)
        first_written, truncated = 0, False
        for k, (t, d) in enumerate(ents):
            i = index + 1 + k
            if i <= len(self.ents):
                if self.ents[i - 1][0] != t:
                    del self.ents[i - 1:]
                    truncated = True
                    self.ents.append((t, d))
                    first_written = first_written or i
            else:
                self.ents.append((t, d))
                first_written = first_written or i
        return lastnewi, True, first_written, truncated


def frame_record(rec: dict) -> bytes:
    """one WAL record: u32 length, u32 CRC-32 of the payload, JSON payload"""
    b = json.dumps(rec, separators=(",", ":")).encode()
    return struct.pack("<II", len(b), zlib.crc32(b)) + b


def scan_records(path: str):
    """-> (records, valid_end): every record up to the first torn / garbled one, and the byte offset where the
    valid prefix ends (a length that outruns the file, a CRC mismatch, or half a header all end the scan)."""
    recs = []
    if not os.path.exists(path):
        return recs, 0
    with open(path, "rb") as f:
        buf = f.read()
    off = 0
    while off + 8 <= len(buf):
        n, crc = struct.unpack_from("<II", buf, off)
        if off + 8 + n > len(buf):
            break  # torn tail record
        body = buf[off + 8: off + 8 + n]
        if zlib.crc32(body) != crc:
            break  # a record whose bytes never all reached the disk: everything before it stands
        try:
            recs.append(json.loads(body))
        except ValueError:
            break
        off += 8 + n
    return recs, off


def repair_tail(path: str) -> int:
    """truncate the file to its valid prefix and make that durable; returns the number of bytes cut"""
    if not os.path.exists(path):
        return 0
    _, end = scan_records(path)
    size = os.path.getsize(path)
    if end < size:
        with open(path, "r+b") as f:
            f.truncate(end)
            f.flush()
            os.fsync(f.fileno())
    return size - end


class Wal:
    """A minimal write-ahead log standing in for etcd `wal` (reference raft.go:100-124,228): a directory
    `raftsql-<id>` with one append-only file of length-prefixed JSON records
    ({"hs": [term, vote, commit]} / {"e": [index, term, data-hex]} / {"t": index} truncate-after)."""

    def __init__(self, dirname: str):
        self.dir = dirname
        self.path = os.path.join(dirname, "wal.log")
        self.f = None

    @staticmethod
    def exist(dirname: str) -> bool:  # wal.Exist (raft.go:100,145)
        return os.path.exists(os.path.join(dirname, "wal.log"))

    def open(self):
        """Open for appending.  A torn or garbled tail (a crash mid-append) is CUT OFF first, as etcd's wal repairs
        its tail before reuse: appending behind the garbage would hide every later record from the next replay."""
        os.makedirs(self.dir, mode=0o750, exist_ok=True)  # raft.go:101
        repair_tail(self.path)
        self.f = open(self.path, "ab")

    def read_all(self):
        """-> (hardstate or None, entries [(term, data)])  (wal.ReadAll, raft.go:124)"""
        hs, ents = None, []
        for rec in scan_records(self.path)[0]:
            if "hs" in rec:
                hs = tuple(rec["hs"])
            elif "e" in rec:
                i, t, d = rec["e"]
                del ents[i - 1:]
                ents.append((t, bytes.fromhex(d)))
            elif "t" in rec:
                del ents[rec["t"]:]
        return hs, ents

    def _put(self, rec):
        self.f.write(frame_record(rec))

    def save(self, hardstate, new_entries, first_index, truncate_after=None):  # wal.Save (raft.go:228)
        if truncate_after is not None:
            self._put({"t": truncate_after})
        for k, (t, d) in enumerate(new_entries):
            self._put({"e": [first_index + k, t, d.hex()]})
        if hardstate is not None:
            self._put({"hs": list(hardstate)})
        self.f.flush()
        os.fsync(self.f.fileno())

    def close(self):
        if self.f:
            self.f.close()
            self.f = None


class LocalTransport:
    """In-process stand-in for rafthttp.Transport (reference raft.go:170-184,230,259): Send() drops a
    message into the destination node's mailbox; unknown / stopped peers lose messages, like a dead TCP peer."""

    def __init__(self):
        self.lock = threading.Lock()
        self.boxes: dict[int, list] = {}

    def register(self, nid: int):
        with self.lock:
            self.boxes[nid] = []

    def unregister(self, nid: int):
        with self.lock:
            self.boxes.pop(nid, None)

    def send(self, msgs):
        with self.lock:
            for m in msgs:
                box = self.boxes.get(m.to)
                if box is not None:
                    box.append(m)

    def drain(self, nid: int):
        with self.lock:
            box = self.boxes.get(nid)
            if box is None:
                return []
            out, box[:] = list(box), []
            return out


class HostNode:
    """One raft node for ONE group (the raftsql shape: G = 1, R = len(peers)) around an engine core."""

    def __init__(self, core, nid: int, npeers: int, transport: LocalTransport, waldir: str | None = None):
        self.core, self.id, self.n = core, nid, npeers
        self.tr = transport
        self.log = Log()
        # a directory name (this node's own WAL) or an object with the Wal interface (a group's view of a shared one)
        self.wal = waldir if hasattr(waldir, "save") else (Wal(waldir) if waldir else None)
        self.pending: list[bytes] = []       # proposals not yet accepted by the state machine
        self.inflight: list[bytes] = []      # proposals posted to the engine this tick
        self.next = [1] * (npeers + 1)       # Progress.Next per peer id (leader only)
        self.behind: set[int] = set()        # peers whose heartbeat response showed Match < lastIndex
        self.backlog: list[Message] = []     # inbound messages deferred to the next tick (one per sender per tick)
        self.applied = 0                     # highest index handed to the commit stream
        self.term = self.vote = self.commit = 0
        self.role, self.lead = F.ROLE_FOLLOWER, 0
        self.replay: list[bytes] = []        # committed payloads recovered from the WAL (published before nil)
        transport.register(nid)

    # -- start / restart (reference raft.go:144-165) ---------------------------------------------------------
    def start(self):
        """replayWAL (raft.go:122-134): rebuild the log, restore HardState, hand back the committed prefix."""
        if self.wal is None:
            return
        old = Wal.exist(self.wal.dir)
        hs, ents = self.wal.read_all() if old else (None, [])
        self.wal.open()
        self.log.ents = list(ents)
        if hs:
            self.term, self.vote, self.commit = hs
        self.commit = min(self.commit, self.log.last_index())
        st = {
            "term": np.array([self.term], np.uint64), "vote": np.array([self.vote], np.uint64),
            "committed": np.array([self.commit], np.uint64),
            "last_index": np.array([self.log.last_index()], np.uint64),
            "last_term": np.array([self.log.last_term()], np.uint64),
        }
        self.core.import_state(st)
        self.replay = [d for (_, d) in self.log.ents[: self.commit] if d]
        self.applied = self.commit

    # -- one iteration of serveChannels (reference raft.go:221-245) --------------------------------------------
    def propose(self, data: bytes):
        self.pending.append(data)

    def step_tick(self) -> list[bytes]:
        """Inbound messages -> engine inbox; proposals; one engine tick; Ready handling.  Returns the payloads
        newly committed by this tick, in log order (what goes to commitC)."""
        app_replies = self.prepare_tick()
        self.core.tick(0)
        return self._ready(app_replies)

    def prepare_tick(self) -> dict:
        """Everything of step_tick that precedes the engine's tick (posts this tick's inbox and proposals).  Split
        out so that a multi-group host can prepare every group, tick the shared engine ONCE, then finish every
        group (raftsql_b200.multipipe)."""
        inbound = self.backlog + self.tr.drain(self.id)
        self.backlog = []
        eng_msgs, app_replies = [], {}
        # the engine inbox holds ONE message per sender per tick (include/mrq.h): a second message from the
        # same peer (say MsgApp then MsgHeartbeat when two of its ticks land in one of ours) waits for the
        # next tick, in order.  At most ONE MsgApp is resolved per tick: a second one (another sender) was
        # matched against a log the first is about to change, so it waits too.
        chosen: dict[int, Message] = {}
        have_app = False
        for m in inbound:
            if m.type == MsgProp:  # a follower forwarded client proposals to us
                self.pending.extend(d for (_, d) in m.entries)
                continue
            if m.frm in chosen or (m.type == F.MSG_APP and have_app):
                self.backlog.append(m)
                continue
            chosen[m.frm] = m
            have_app = have_app or m.type == F.MSG_APP
        # The engine Steps the tick's messages in SENDER order, so a higher-term message from a
        # lower sender id moves the engine to that term before it sees a later sender's MsgApp.  The host must
        # resolve the append against that same effective (term, role), or it would change its log and WAL for
        # an append the engine then drops on the term rule (host log and engine would disagree: ADVICE r1).
        eff_term, eff_role = self.term, self.role
        for frm in sorted(chosen):
            m = chosen[frm]
            if m.term > eff_term:  # Step(): becomeFollower(m.Term, ...)
                eff_term, eff_role = m.term, F.ROLE_FOLLOWER
            if m.type == F.MSG_APP:
                if eff_role == F.ROLE_CANDIDATE and m.term == eff_term and any(
                        c.type == F.MSG_VOTE_RESP and f < frm for f, c in chosen.items()):
                    # votes stepped before it may make us leader within this tick (the append would then be
                    # ignored): let the votes land first, resolve the append next tick
                    self.backlog.insert(0, m)
                    continue
                rec = self._resolve_append(m, app_replies, eff_term, eff_role)
                if rec is not None:
                    eng_msgs.append(rec)
                if eff_role == F.ROLE_CANDIDATE and m.term == eff_term:
                    eff_role = F.ROLE_FOLLOWER  # stepCandidate MsgApp: becomeFollower(Term, From)
                continue
            if m.type == F.MSG_HEARTBEAT and eff_role == F.ROLE_CANDIDATE and m.term == eff_term:
                eff_role = F.ROLE_FOLLOWER
            if m.type == F.MSG_APP_RESP and m.reject and self.role == F.ROLE_LEADER and m.term == self.term:
                # Progress.maybeDecrTo: message-construction state, host side
                self.next[m.frm] = max(1, min(m.index, m.reject_hint + 1))
            if m.type == F.MSG_HEARTBEAT_RESP and self.role == F.ROLE_LEADER and m.term == self.term:
                # stepLeader MsgHeartbeatResp: `if pr.Match < lastIndex { sendAppend }` — resend from Match+1
                self.behind.add(m.frm)
            ty = m.type | (F.MSG_REJECT if m.reject else 0)
            eng_msgs.append((0, m.frm, ty, m.term, m.index, m.logterm, m.commit))
        self.core.post_inbox_delta(eng_msgs, slot=0)
        # node.Propose blocks until there is a leader (upstream node.run: propc is nil while lead == None)
        self.inflight = []
        if self.pending and self.role == F.ROLE_LEADER:
            self.inflight, self.pending = self.pending[:255], self.pending[255:]
            self.core.propose([0], [len(self.inflight)], slot=0)
        elif self.pending and self.lead not in (0, self.id):
            fwd, self.pending = self.pending, []
            self.tr.send([Message(MsgProp, self.lead, self.id, entries=[(0, d) for d in fwd])])
        return app_replies

    def finish_tick(self, app_replies: dict) -> list[bytes]:
        """Everything of step_tick that follows the engine's tick (Ready handling)."""
        return self._ready(app_replies)

    def _resolve_append(self, m: Message, replies: dict, eff_term: int | None = None, eff_role: int | None = None):
        """The log-matching half of handleAppendEntries, host side (include/mrq.h MSG_APP contract), against the
        (term, role) the engine will have when it Steps this message (prepare_tick computes them in sender order)."""
        eff_term = self.term if eff_term is None else eff_term
        eff_role = self.role if eff_role is None else eff_role
        if m.term < eff_term or (eff_role == F.ROLE_LEADER and m.term == eff_term):
            return (0, m.frm, F.MSG_APP, m.term, 0, 0, 0)  # the engine will drop it on the term rule
        if m.index < self.commit:
            replies[m.frm] = Message(F.MSG_APP_RESP, m.frm, self.id, index=self.commit)
            return (0, m.frm, F.MSG_APP, m.term, self.log.last_index(), self.log.last_term(), self.commit)
        lastnewi, ok, first_written, truncated = self.log.maybe_append(m.index, m.logterm, m.commit, m.entries)
        if not ok:
            replies[m.frm] = Message(F.MSG_APP_RESP, m.frm, self.id, index=m.index, reject=True,
                                     reject_hint=self.log.last_index())
            return (0, m.frm, F.MSG_APP | F.MSG_REJECT, m.term, 0, 0, 0)
        # persist what changed before acknowledging (wal.Save precedes transport.Send, raft.go:228-230)
        if self.wal is not None and first_written:
            self.wal.save(None, self.log.ents[first_written - 1:], first_written,
                          truncate_after=(first_written - 1) if truncated else None)
        replies[m.frm] = Message(F.MSG_APP_RESP, m.frm, self.id, index=lastnewi)
        return (0, m.frm, F.MSG_APP, m.term, self.log.last_index(), self.log.last_term(), min(m.commit, lastnewi))

    def _ready(self, app_replies: dict) -> list[bytes]:
        s = self.core.export_state(("term", "vote", "committed", "last_index", "last_term", "term_start", "match",
                                    "role", "lead"))
        out = int(self.core.sync_out()[0])
        was_leader = self.role == F.ROLE_LEADER
        term, vote, commit = int(s["term"][0]), int(s["vote"][0]), int(s["committed"][0])
        role, lead, last = int(s["role"][0]), int(s["lead"][0]), int(s["last_index"][0])
        hs_changed = (term, vote, commit) != (self.term, self.vote, self.commit)
        self.term, self.vote, self.role, self.lead = term, vote, role, lead
        msgs: list[Message] = []
        # entries the engine appended as leader: the empty entry of a new term, then accepted proposals
        new_entries = []
        if role == F.ROLE_LEADER and last > self.log.last_index():
            n_new = last - self.log.last_index()
            if out & F.OUT_BECAME_LEADER:
                new_entries.append((term, b""))
                n_new -= 1
            accepted, rest = self.inflight[:n_new], self.inflight[n_new:]
            new_entries.extend((term, d) for d in accepted)
            self.pending = rest + self.pending
            first = self.log.last_index() + 1
            self.log.ents.extend(new_entries)
            if out & F.OUT_BECAME_LEADER:
                self.next = [self.log.last_index() - len(new_entries) + 1] * (self.n + 1)  # reset(): Next = lastIndex+1
        elif self.inflight:  # stepped down before the proposals were applied: nothing was appended
            self.pending = self.inflight + self.pending
        self.inflight = []
        if self.wal is not None and (new_entries or hs_changed):
            self.wal.save((term, vote, commit), new_entries, (first if new_entries else 0))
        # --- Ready.Messages, rebuilt from the out word (include/mrq.h MRQ_OUT_*) -----------------------------
        if out & F.OUT_CAMPAIGN:
            for p in self._peers():
                msgs.append(Message(F.MSG_VOTE, p, self.id, term=term, index=last, logterm=int(s["last_term"][0])))
        for p in self._peers():
            rep = (out >> (F.OUT_VOTE_REPLY_SHIFT + 2 * (p - 1))) & 3
            if rep:
                msgs.append(Message(F.MSG_VOTE_RESP, p, self.id, term=term, reject=(rep == 2)))
            if (out >> (F.OUT_ACK_REPLY_SHIFT + (p - 1))) & 1:
                r = app_replies.get(p) or Message(F.MSG_HEARTBEAT_RESP, p, self.id)
                r.term = term
                msgs.append(r)
        if role == F.ROLE_LEADER and (out & (F.OUT_BCAST_APPEND | F.OUT_BCAST_HEARTBEAT | F.OUT_BECAME_LEADER)):
            match = s["match"][:, 0]
            for p in self._peers():
                if p in self.behind and int(match[p - 1]) < self.log.last_index():
                    self.next[p] = int(match[p - 1]) + 1
                nxt = max(self.next[p], int(match[p - 1]) + 1)
                if nxt <= self.log.last_index():  # sendAppend
                    ents = self.log.slice(nxt)
                    msgs.append(Message(F.MSG_APP, p, self.id, term=term, index=nxt - 1, logterm=self.log.term(nxt - 1),
                                        entries=ents, commit=commit))
                    self.next[p] = nxt + len(ents)  # optimistic, like ProgressStateReplicate
                elif out & F.OUT_BCAST_HEARTBEAT:
                    msgs.append(Message(F.MSG_HEARTBEAT, p, self.id, term=term, commit=min(int(match[p - 1]), commit)))
        self.behind.clear()
        if was_leader and role != F.ROLE_LEADER:
            self.next = [1] * (self.n + 1)
        self.tr.send(msgs)  # transport.Send(rd.Messages) (raft.go:230)
        # --- publish (raft.go:82-96, but gated on commit) ------------------------------------------------------
        self.commit = commit
        published = []
        while self.applied < min(self.commit, self.log.last_index()):
            self.applied += 1
            d = self.log.ents[self.applied - 1][1]
            if d:  # "ignore conf changes and empty messages" (raft.go:84-86)
                published.append(d)
        # the engine tracks (lastIndex, lastTerm) of the log the host keeps: after Ready they must agree, or every
        # later vote / commit decision is made about a log that does not exist (fatal, like the reference's log.Fatalf)
        if (last, int(s["last_term"][0])) != (self.log.last_index(), self.log.last_term()):
            raise RuntimeError(f"node {self.id}: host log ({self.log.last_index()}, t{self.log.last_term()}) and engine "
                               f"({last}, t{int(s['last_term'][0])}) diverged")
        return published

    def _peers(self):
        return [p for p in range(1, self.n + 1) if p != self.id]

    def stop(self):
        self.tr.unregister(self.id)
        if self.wal is not None:
            self.wal.close()
