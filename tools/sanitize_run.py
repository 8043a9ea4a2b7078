#!/usr/bin/env python
"""A small workload that touches every kernel of libmrq.so, for compute-sanitizer (GPU box):

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_run.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_run.py

Sizes are deliberately ragged (G not a multiple of the CTA / tile sizes) so that tail handling is exercised.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raftsql_b200 import Engine, empty_inbox, preset_trace  # noqa: E402
from raftsql_b200 import _ffi as F  # noqa: E402
from raftsql_b200.packed import pack_inbox, pack_inbox16  # noqa: E402


def main():
    G, R = 1543, 5
    p = preset_trace(5)
    for mode in (0, 1, 2):
        eng = Engine(G, R, seed=11, inbox_slots=3)
        eng.set_tick_mode(mode)
        for t in range(30):
            eng.gen_trace(p, t, slot=t % 3)
            eng.tick(t % 3)
        eng.tick_many([0, 1, 2, 0, 1])      # graph replay path (small shard => auto graphs)
        eng.tick_many([0, 1, 2, 0, 1])
        eng.tick_idle(3)
        st = eng.export_state()
        eng.import_state(st)
        for variant in (0, 1, 2):
            eng.set_quorum_variant(variant)
            eng.quorum_commit()
        lead = np.nonzero(st["role"] == 2)[0].astype(np.uint64)
        if len(lead):
            eng.match_update(lead, np.full(len(lead), 1 + (st["self_id"][lead[0]] % R), np.uint8), st["last_index"][lead])
        ib = eng.read_inbox(0)
        base_i = np.where(st["last_index"] > 100, st["last_index"] - np.uint64(100), 0).astype(np.uint64)
        eng.set_packed_base(base_i, st["term"])
        for packer in (pack_inbox, pack_inbox16):
            w, p8, wide = packer(ib, base_i, st["term"])
            eng.post_inbox_packed(w, p8, wide, slot=1)
            eng.tick(1)
        msgs = [(int(g), 1 + (int(g) % R), F.MSG_APP_RESP, int(st["term"][g]), int(st["last_index"][g]), 0, 0) for g in range(0, G, 7)]
        eng.post_inbox_delta(msgs, slot=2)
        eng.propose(np.arange(0, G, 5, dtype=np.uint64), np.ones(len(range(0, G, 5)), np.uint32), slot=2)
        eng.tick(2)
        eng.post_inbox_dense(empty_inbox(G, R), slot=0)
        eng.tick(0)
        eng.sync_commit_deltas()
        eng.sync_commits(want_role=True, want_term=True)
        eng.sync_out()
        eng.export_next()
        c = eng.counters()
        assert c["errors"] == 0
        eng.close()
    print("sanitize_run: ok")


if __name__ == "__main__":
    main()
