"""TEST-ONLY stand-in for the GPU engine behind raftsql_b200.hostnode / raftpipe: the CPU oracle wearing the
Engine's tick protocol, so the host-side logic (log matching, message construction, WAL, channel protocol,
SQLite apply loop, HTTP surface) is exercised by the CPU suite too.  Lives under tests/ on purpose: nothing in
the product package may touch the oracle (the same plumbing tests run against the real Engine under
`-m gpu`)."""
import numpy as np

import oracle


class OracleCore:
    def __init__(self, npeers: int, nid: int, seed: int = 0, election_tick: int = 10, heartbeat_tick: int = 1, n_groups: int = 1, **_):
        self.G, self.R = n_groups, npeers
        self.o = oracle.Oracle(n_groups, npeers, self_id=nid, seed=seed or (0x5EED + nid), election_tick=election_tick,
                               heartbeat_tick=heartbeat_tick)
        self.ib = oracle.empty_inbox(n_groups, npeers)

    def import_state(self, s):
        self.o.import_state({k: np.ascontiguousarray(v) for k, v in s.items()})

    def post_inbox_delta(self, msgs, slot=0, accumulate=False):
        if not accumulate:
            self.ib = oracle.empty_inbox(self.G, self.R)
        for g, frm, ty, term, index, logterm, commit in msgs:
            r = frm - 1
            self.ib["type"][r, g], self.ib["term"][r, g], self.ib["index"][r, g] = ty, term, index
            self.ib["logterm"][r, g], self.ib["commit"][r, g] = logterm, commit

    def propose(self, groups, counts, slot=0):
        for g, c in zip(groups, counts):
            self.ib["prop_count"][g] += c

    def tick(self, slot=0):
        self.o.tick(self.ib)
        self.ib = oracle.empty_inbox(self.G, self.R)

    def export_state(self, columns=None):
        return self.o.export()

    def sync_out(self):
        return self.o.export()["out"]

    def close(self):
        self.o.close()


def make_oracle_core(npeers, nid, **kw):
    return OracleCore(npeers, nid, **kw)


def make_oracle_multicore(npeers, nid, n_groups, **kw):
    """the multi-group core of raftsql_b200.multipipe (one engine of G groups), CPU checker version"""
    return OracleCore(npeers, nid, n_groups=n_groups, **kw)
