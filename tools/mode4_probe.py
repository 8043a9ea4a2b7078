#!/usr/bin/env python
"""Quick probe (development): µs per tick of tick mode 4 at 1,048,576 x 5 — per-tick launches vs one launch per K ticks,
write-through on/off — next to mode 0 (wide inbox) and mode 3 (byte inbox on wide state).  Not a bench: no JSON contract."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from raftsql_b200 import Engine, preset_trace  # noqa: E402
from raftsql_b200.packed import Pack8  # noqa: E402

G = int(os.environ.get("PROBE_G", 1 << 20))
R, K, W = 5, 20, 5
NS = K + W
st0 = bench.steady_state(G, R, 0, bench.SEED)
p = preset_trace(3)
eng = Engine(G, R, seed=bench.SEED, inbox_slots=NS)
eng.import_state(st0)
t0 = time.time()
for t in range(NS):
    eng.gen_trace(p, t, slot=t)
    eng.tick(t)
eng.synchronize()
base0 = (st0["last_index"] - np.uint64(40)).astype(np.uint64)
pk = Pack8(st0["self_id"], base0, st0["term"], R)
frames = [pk.frame(eng.read_inbox(t)) for t in range(NS)]
print(f"trace + frames in {time.time() - t0:.1f} s; escapes {sum(len(f[2]) for f in frames)}", flush=True)


def rewind(mode):
    eng.set_tick_mode(0)
    eng.import_state(st0)
    eng.tick_count = 0
    eng.set_tick_mode(mode)
    if mode >= 3:
        eng.set_packed_base(base0, st0["term"])


def timed(label, mode, graph, wt=1, reps=5):
    res = []
    for _ in range(reps):
        rewind(mode)
        eng.set_graph_mode(graph)
        eng.set_write_through(wt)
        eng.tick_many(list(range(W)))
        eng.synchronize()
        eng.timer_start()
        eng.tick_many([W + k for k in range(K)])
        res.append(eng.timer_stop() / K * 1e3)
    c = eng.sync_commits()
    print(f"{label:58s} us/tick {np.median(res):7.2f}  (reps {' '.join(f'{x:.2f}' for x in res)})  checksum {int(c.sum() % (1 << 32))}", flush=True)
    return c


ref = timed("mode 0: wide inbox, fast+slow launches per tick", 0, 0)
eng.set_tick_mode(3)
for t, (w8, p8, wide8) in enumerate(frames):
    eng.post_inbox_packed(w8, p8, wide8, slot=t, keep=True)
c3 = timed("mode 3: byte inbox on wide state, per-tick launches", 3, 0)
c4a = timed("mode 4: compact + bytes, per-tick launches", 4, 0)
c4b = timed("mode 4: one launch per 20 ticks, write-through", 4, 2, 1)
c4c = timed("mode 4: one launch per 20 ticks, write-back at the end", 4, 2, 0)
for name, c in (("mode3", c3), ("mode4 per-tick", c4a), ("mode4 batch wt", c4b), ("mode4 batch", c4c)):
    print(name, "equals mode 0:", bool(np.array_equal(c, ref)))
print(eng.counters())
