#!/usr/bin/env python
"""Summarise ncu captures from gpurun_out/ into profiles/ (run here: ncu reads .ncu-rep without a GPU).

    python tools/ncu_summary.py r01 gpurun_out/launches_r01.csv gpurun_out/prof_tickfast_r01.ncu-rep ...

Writes profiles/<round>_launches.md (share of the step per kernel, from the launch list) and
profiles/<round>_<report>.md (key raw metrics + stall breakdown + top stall sites) per .ncu-rep.
"""
import csv
import io
import os
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("NCU_SUMMARY_OUT") or os.path.join(ROOT, "profiles")  # (on the GPU box: a directory under gpurun_out/)

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
       "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
       "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
       "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches(round_id, path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    order = []
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("mrq::", "")
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        agg[name][0] += 1
        agg[name][1] += v
        order.append((name, v))
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{round_id}_launches.md"), "w") as f:
        f.write(f"# ncu launch list, round {round_id}\n\n"
                f"Source: `{os.path.basename(path)}` — `ncu --metrics gpu__time_duration.sum --clock-control none` over "
                "`MRQ_BENCH_FAST=1 python bench.py --steps 10 --warmup 3` (cold-cache, serialised: compare SHARES).\n\n"
                "| kernel | launches | total ns | share | mean ns |\n|---|---:|---:|---:|---:|\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {t:.0f} | {100 * t / total:.1f}% | {t / n:.0f} |\n")
        # the timed step itself: tick_fast + tick_slow launches only
        step = {k: v for k, v in agg.items() if k.startswith("tick_")}
        st = sum(v[1] for v in step.values())
        if st:
            f.write("\nWithin one tick (the timed step = `tick_fast_kernel` + `tick_slow_kernel`):\n\n")
            for name, (n, t) in sorted(step.items(), key=lambda kv: -kv[1][1]):
                f.write(f"- `{name}`: {100 * t / st:.1f}% of the step ({t / n:.0f} ns mean)\n")


def report(round_id, path):
    base = os.path.splitext(os.path.basename(path))[0]
    raw = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    lines = [f"# {base}\n", f"Source: `{os.path.basename(path)}` (`ncu --set full --clock-control none --import-source on`).\n"]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")]
        lines.append(f"\n## `{name}`\n\n| metric | value |\n|---|---|\n")
        for m in RAW:
            if m in hdr:
                lines.append(f"| {m} | {row[hdr.index(m)]} {units[hdr.index(m)]} |\n")
        try:
            rd = float(row[hdr.index("dram__bytes_read.sum")])
            wr = float(row[hdr.index("dram__bytes_write.sum")])
            u_r, u_w = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd * mult[u_r] + wr * mult[u_w]
            dur = float(row[hdr.index("gpu__time_duration.sum")]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3}[
                units[hdr.index("gpu__time_duration.sum")]]
            lines.append(f"\nDRAM traffic (read + write) = {tot / 1e6:.1f} MB per launch; {tot / dur / 1e12:.2f} TB/s over the "
                         "kernel's duration (cold cache, isolated).\n")
        except Exception:
            pass
    src = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "source", "--csv"]))))
    if len(src) > 2:
        h = src[1]
        iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        data, seen = [], set()
        for r in src[2:]:
            if len(r) > iE and r[iE].isdigit():
                if r[h.index("Address")] in seen:
                    break
                seen.add(r[h.index("Address")])
                data.append(r)
        tot = sum(int(r[iE]) for r in data)
        ops, st = Counter(), Counter()
        for r in data:
            s = r[iS]
            op = s.split()[1] if s.startswith("@") else s.split()[0]
            ops[op.split(".")[0]] += int(r[iE])
            for i in stall:
                try:
                    st[h[i]] += int(r[i])
                except ValueError:
                    pass
        lines.append(f"\n## SASS (first captured launch)\n\n{len(data)} SASS instructions; {tot} warp-instructions executed.\n\n")
        lines.append("Opcode mix: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(10)) + "\n\n")
        ts = sum(st.values()) or 1
        lines.append("Warp stalls (sampled): " + ", ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in st.most_common(6)) + "\n\n")
        proof = sorted({op for op in ops if op in ("UBLKCP", "UTMALDG", "SYNCS", "LDG", "STG", "VIMNMX", "REDUX", "VOTE", "ATOMG", "RED")})
        lines.append("Notable SASS present: " + ", ".join(proof) + "\n\nTop stall sites:\n\n```\n")
        for r in sorted(data, key=lambda r: -int(r[iSm] or 0))[:8]:
            lines.append(f"{r[iSm]:>6} samples  {r[iS][:110]}\n")
        lines.append("```\n")
    with open(os.path.join(OUT, f"{round_id}_{base}.md"), "w") as f:
        f.writelines(lines)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    rid = sys.argv[1]
    for p in sys.argv[2:]:
        if p.endswith(".csv"):
            launches(rid, p)
        else:
            report(rid, p)
        print("summarised", p)
