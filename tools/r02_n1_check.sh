#!/bin/bash
# One N = 1 GPU call: the full -m gpu suite, the K3 shape explorer, the default bench (summarised).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r02_n1_check.sh <tag>'
set -u
OUT=gpurun_out/r02_${1:-n1}
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/gpu_suite.log" 2>&1
echo "suite exit $?"
tail -6 "$OUT/gpu_suite.log"
grep -E "^(FAILED|ERROR)" "$OUT/gpu_suite.log"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/k3_explore tools/k3_explore.cu 2>/dev/null \
  && timeout 300 /tmp/k3_explore "$(python -c "import json;print(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])" 2>/dev/null || echo 6583.5)" > "$OUT/k3_explore.txt" 2>&1
cat "$OUT/k3_explore.txt"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?"
python - "$OUT/bench.json" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value", round(l["value"]), "us", round(l["ms_per_step"] * 1e3, 2), "frac", round(l["roofline"]["frac"], 3), "clocks", l["clocks"])
    for k, v in l.get("variants", {}).items():
        print("  ", k, round(v["us_per_tick"], 2), "us  frac", round(v["frac"], 3))
    e = l["e2e"]
    print("e2e", round(e["value"]), "pack_us", round(e["pack_us_per_tick"]), "equal", e["equals_wide_form"], "| preencoded", round(e["preencoded"]["value"]),
          "h2d GB/s", round(e["preencoded"]["h2d_GBps_per_gpu"], 1))
    print("cpu", l["cpu_baseline"]["value"], "| k3", {k: round(v["us_per_launch"], 2) for k, v in l["roofline_quorum_kernel"]["variants"].items()})
except Exception as ex:
    print("could not read the bench line:", ex)
PY
tail -3 "$OUT/bench.err"
