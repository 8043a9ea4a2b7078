"""raftdb — the replicated SQLite state machine above the seam, mirroring reference db.go:13-167 so the
config-1 plumbing (3-node cluster, CREATE/INSERT/SELECT) can be exercised without a Go toolchain.
This layer is OUT of the hot path; it exists to show the seam is a drop-in.
"""
from __future__ import annotations

import os
import sqlite3
import threading

from .raftpipe import Chan, RaftPipe


def isSelect(query: str) -> bool:
    """reference db.go:98-104: trim spaces only, split on a single space, first token case-insensitively SELECT."""
    tokens = query.strip(" ").split(" ")
    return len(tokens) > 0 and tokens[0].upper() == "SELECT"


class RaftDB:
    def __init__(self, path: str, rp: RaftPipe, commitC: Chan | None = None):
        # "database is entirely replayed from the raft log until snapshots are supported" (db.go:27-29)
        try:
            os.remove(path)
        except FileNotFoundError:
            pass
        self.db = sqlite3.connect(path, check_same_thread=False, isolation_level=None)
        self.rp = rp
        self.dbmu = threading.RLock()
        self.mu = threading.Lock()
        self.q2cb: dict[str, list[Chan]] = {}
        self.listener = commitC
        self.fatal = None
        self.readCommits()  # synchronous replay until the None sentinel (db.go:40)
        self._th = threading.Thread(target=self.readCommits, daemon=True)  # live stream (db.go:41)
        self._th.start()

    def readCommits(self):
        """reference db.go:45-96"""
        for q in self.rp.CommitC:
            if q is None:
                if self.listener is not None:
                    self.listener.send(None)
                return
            with self.dbmu:
                try:
                    self.db.execute(q)
                    err = None
                except sqlite3.Error as ex:
                    err = ex
            if self.listener is not None:
                self.listener.send(q)
            with self.mu:
                cbcs = self.q2cb.get(q)
                if not cbcs:
                    continue  # either a replay or from another node (db.go:64-69)
                cur = cbcs.pop(0)  # exact-text FIFO match (db.go:70-75)
                if not cbcs:
                    del self.q2cb[q]
            cur.send(err)
            cur.close()
        err, ok = self.rp.ErrorC.recv()
        if ok:  # db.go:83-95: fail every waiter; the reference then log.Fatal()s
            with self.mu:
                for v in self.q2cb.values():
                    for c in v:
                        c.send(err)
                        c.close()
                self.q2cb = {}
            self.db.close()
            self.fatal = err

    def Propose(self, query: str) -> Chan:
        """reference db.go:106-121: returns a channel that yields the apply error (None on success)."""
        errc = Chan(buffered=1)
        if isSelect(query):
            errc.send(ValueError("expected non-SELECT"))
            return errc
        with self.mu:
            self.q2cb.setdefault(query, []).append(errc)
        self.rp.ProposeC.send(query)
        return errc

    def Query(self, query: str) -> str:
        """reference db.go:123-157: SELECT only; rows rendered |c1|c2|...|\\n, NULL as empty.  Raises on error."""
        if not isSelect(query):
            raise ValueError("expected SELECT")
        with self.dbmu:
            cur = self.db.execute(query)
            rows = cur.fetchall()
        ret = ""
        for row in rows:
            for v in row:
                if v is None:
                    s = ""
                elif isinstance(v, bytes):
                    s = v.decode(errors="replace")
                else:
                    s = str(v)
                ret += "|" + s
            ret += "|\n"
        return ret

    def Close(self):
        """reference db.go:159-167"""
        with self.mu:
            if self.q2cb:
                raise RuntimeError("Closing db with outstanding callbacks")  # log.Fatalf in the reference
        err = self.rp.Close()
        self._th.join(timeout=10)
        try:
            self.db.close()
        except Exception:
            pass
        return err


def NewDB(path: str, rp: RaftPipe) -> RaftDB:  # db.go:22-24
    return RaftDB(path, rp, None)


def NewDBListen(path: str, rp: RaftPipe, commitC: Chan | None) -> RaftDB:  # db.go:26-43
    return RaftDB(path, rp, commitC)
