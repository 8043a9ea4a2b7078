#!/usr/bin/env python
"""Print the key figures of a bench.py JSON line (development helper)."""
import json
import sys

l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(l["value"]), "us", round(l["ms_per_step"] * 1e3, 2), "frac", round(l["roofline"]["frac"], 3), "traffic", l["roofline"]["traffic"],
      "clocks", l["clocks"])
for k, v in (l.get("variants") or {}).items():
    print("  ", k, round(v["us_per_tick"], 2), "us  frac", round(v["frac"], 3))
e = l.get("e2e")
if e:
    print("e2e", round(e["value"]), "pack_us", round(e["pack_us_per_tick"]), "equal", e["equals_wide_form"], "| pre", round(e["preencoded"]["value"]),
          "h2d GB/s", round(e["preencoded"]["h2d_GBps_per_gpu"], 1))
q = l.get("roofline_quorum_kernel")
if q:
    print("k3", {k: round(v["us_per_launch"], 2) for k, v in q["variants"].items()}, "traffic", q["traffic"])
if l.get("cpu_baseline"):
    print("cpu", l["cpu_baseline"]["value"])
