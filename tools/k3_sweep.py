#!/usr/bin/env python
"""Sweep the standalone quorum kernel over problem sizes and kernel forms (development tool, GPU only).

Separates the fixed cost of a launch (ramp-up, tail, inter-launch gap) from the steady-state HBM rate:
time(G) ~= t_fixed + (8R+16) * G / BW.  Every timed launch reads a column set that was never touched since
the last L2 flush.  Prints one JSON line per (G, variant).
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raftsql_b200 import Engine  # noqa: E402

R = 5
NAMES = {0: "ldg256", 1: "tma_bulk", 2: "ldg128"}


def main():
    eng = Engine(64, R)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for logG in (18, 19, 20, 21, 22, 23):
        G = 1 << logG
        nsets = max(6, min(24, (6 << 30) // (56 * G)))
        sets = []
        for _ in range(nsets):
            li = torch.randint(2 ** 20, 2 ** 40, (G,), generator=gen, device="cuda", dtype=torch.int64)
            lag = torch.randint(0, 12, (R, G), generator=gen, device="cuda", dtype=torch.int64)
            sets.append(((li.unsqueeze(0) - lag).contiguous(), (li - 40).contiguous(), (li - 45).contiguous()))
        for variant in (2, 0, 1):
            best = None
            for rep in range(3):
                for m, c, g in sets:
                    c.copy_(g + 5)
                flush.fill_(rep)
                torch.cuda.synchronize()
                eng.timer_start()
                for m, c, g in sets:
                    eng.quorum_commit_ext(m.data_ptr(), c.data_ptr(), g.data_ptr(), G, G, variant)
                ms = eng.timer_stop()
                per = ms / nsets
                best = per if best is None else min(best, per)
            print(json.dumps({"G": G, "variant": NAMES[variant], "us_per_launch": round(best * 1e3, 2),
                              "GBps": round((8 * R + 16) * G / (best * 1e-3) / 1e9, 1), "sets": nsets}), flush=True)
        del sets
        torch.cuda.empty_cache()
    eng.close()


if __name__ == "__main__":
    main()
