#!/usr/bin/env python
"""Host-only probe (development): time of one mrq_pack8 frame at 1,048,576 x 5 for the thread count in MRQ_HOST_THREADS."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from raftsql_b200.packed import Pack8  # noqa: E402

G, R = 1 << 20, 5
st0 = bench.steady_state(G, R, 0, bench.SEED)
ibs = bench.host_inboxes_from_oracle(G, R, st0, 4)
base0 = (st0["last_index"] - np.uint64(40)).astype(np.uint64)
w = np.zeros((R - 1, G), np.uint8)
p8 = np.zeros(G, np.uint8)
best = []
for rep in range(4):
    pk = Pack8(st0["self_id"], base0, st0["term"], R)
    ts = []
    for ib in ibs:
        t = time.perf_counter()
        pk.frame(ib, word_out=w, prop8_out=p8)
        ts.append(time.perf_counter() - t)
    best.append(min(ts[1:]))
print("threads", os.environ.get("MRQ_HOST_THREADS", "default"), "best us per frame", round(min(best) * 1e6), "median", round(sorted(best)[len(best) // 2] * 1e6))
