#!/usr/bin/env python
"""profiles/r02_traffic.json from the raw ncu pages of tools/r02_prof.sh (gpurun_out/r02_prof/*.raw.csv).

Each entry: DRAM bytes read / written and duration of one captured launch (ncu --set full, cold cache, isolated), and the
static SASS instruction count of that kernel in the raftsql_b200/libmrq.so that was profiled — bench.py refuses to quote
a traffic figure when the library it loads has a different count for the kernel (a capture of another build)."""
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_prof")
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
# report file -> (key in the JSON, substring of the mangled kernel name, note)
WHAT = {
    "prof_tick4_pertick_r02": ("tick_fast4_kernel<5>", "tick_fast4_kernelILi5", "tick mode 4, one launch pair per tick, L2 hints on"),
    "prof_tick4_pertick_l2off_r02": ("tick_fast4_kernel<5> (l2 hints off)", "tick_fast4_kernelILi5", "tick mode 4, per tick, mrq_set_l2_policy(0)"),
    "prof_tick4_batched_r02": ("tick_fast4_kernel<5> (20 ticks per launch)", "tick_fast4_kernelILi5", "tick mode 4, ONE launch for 20 ticks, write-through"),
    "prof_tickslow4_r02": ("tick_slow4_kernel<5>", "tick_slow4_kernelILi5", "general path of mode 4 (near-empty list on this trace)"),
    "prof_tick3_r02": ("tick_fast8_kernel<5>", "tick_fast8_kernelILi5", "tick mode 3 (byte inbox, wide state)"),
    "prof_tick0_r02": ("tick_fast_kernel<5>", "tick_fast_kernelILi5", "tick mode 0 (wide inbox, wide state): round 1's kernel"),
    "prof_k3_ldg256_r02": ("quorum_kernel_ldg256<5>", "quorum_kernel_ldg256ILi5", "K3, 256-bit loads, on the engine's own columns"),
    "prof_k3_tma_r02": ("quorum_kernel_tma<5>", "quorum_kernel_tmaILi5", "K3, TMA bulk copies"),
    "prof_k3_ldg128_r02": ("quorum_kernel_ldg<5>", "quorum_kernel_ldgILi5", "K3, 128-bit loads"),
}
out = {"_note": "dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum of ONE launch per kernel from tools/r02_prof.sh "
                "(ncu --set full --clock-control none, cold cache, kernel in isolation); sass_instructions = static SASS count of the kernel "
                "in the profiled libmrq.so (bench.py checks it against the library it loads). Stores still dirty in L2 when a kernel "
                "ends are not in its dram write figure."}
for path in sorted(glob.glob(os.path.join(SRC, "*.raw.csv"))):
    base = os.path.basename(path)[: -len(".raw.csv")]
    if base not in WHAT:
        continue
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        continue
    hdr, units, row = rows[0], rows[1], rows[-1]  # the last captured launch: steady state
    def val(name, table):
        i = hdr.index(name)
        return float(row[i].replace(",", "")) * table[units[i]]
    key, mangled, note = WHAT[base]
    out[key] = {"dram_read_bytes": int(val("dram__bytes_read.sum", MULT)), "dram_write_bytes": int(val("dram__bytes_write.sum", MULT)),
                "duration_us": round(val("gpu__time_duration.sum", TIME) * 1e6, 3),
                "l2_hit_rate_pct": round(float(row[hdr.index("lts__t_sector_hit_rate.pct")]), 1),
                "warp_instructions": int(float(row[hdr.index("smsp__inst_executed.sum")].replace(",", ""))),
                "registers": int(float(row[hdr.index("launch__registers_per_thread")])),
                "sass_instructions": bench.sass_instruction_count(mangled), "what": note, "source": f"profiles/r02_{base}.md"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
