#!/usr/bin/env python
"""Profiling driver (development): runs ONE scenario at 1,048,576 x 5 so that ncu can capture its kernels.

    python tools/prof_r02.py tick4 [l2off]   tick mode 4, one launch pair per tick, 10 ticks
    python tools/prof_r02.py tick4batch      tick mode 4, ONE launch pair for 20 ticks (mrq_tick_many)
    python tools/prof_r02.py tick3 | tick0   tick modes 3 / 0, per-tick launches, 10 ticks
    python tools/prof_r02.py k3 <variant>    the standalone quorum kernel on the engine's own (freshly imported) columns
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from raftsql_b200 import Engine, preset_trace  # noqa: E402
from raftsql_b200.packed import Pack8  # noqa: E402

what = sys.argv[1]
G, R, NS = 1 << 20, 5, 26
st0 = bench.steady_state(G, R, 0, bench.SEED)
eng = Engine(G, R, seed=bench.SEED, inbox_slots=NS if what != "k3" else 2)
eng.import_state(st0)
if what == "k3":
    eng.set_quorum_variant(int(sys.argv[2]))
    for _ in range(3):  # every launch on freshly imported columns: every group's commit index advances
        eng.import_state(st0)
        eng.quorum_commit()
        eng.synchronize()
    print("k3 done", eng.counters()["commits_advanced"])
    sys.exit(0)
p = preset_trace(3)
for t in range(NS):
    eng.gen_trace(p, t, slot=t)
    eng.tick(t)
eng.synchronize()
mode = {"tick4": 4, "tick4batch": 4, "tick3": 3, "tick0": 0}[what]
base0 = (st0["last_index"] - np.uint64(40)).astype(np.uint64)
if mode >= 3:
    pk = Pack8(st0["self_id"], base0, st0["term"], R)
    frames = [pk.frame(eng.read_inbox(t)) for t in range(NS)]
    eng.set_tick_mode(mode)
    for t, (w8, p8, wide8) in enumerate(frames):
        eng.post_inbox_packed(w8, p8, wide8, slot=t, keep=True)
eng.set_tick_mode(0)
eng.import_state(st0)
eng.tick_count = 0
eng.set_tick_mode(mode)
if mode >= 3:
    eng.set_packed_base(base0, st0["term"])
if "l2off" in sys.argv:
    eng.set_l2_policy(0)
if what == "tick4batch":
    eng.set_graph_mode(2)
    eng.tick_many(list(range(5)))          # launch pair 1: the first tick compacts every group (general path)
    eng.tick_many(list(range(5, 25)))      # launch pair 2: 20 steady-state ticks in one launch
else:
    eng.set_graph_mode(0)
    for t in range(12):
        eng.tick(t)
eng.synchronize()
print(what, "done", eng.counters())
