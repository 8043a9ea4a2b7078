#!/bin/bash
# The first GPU call of round 2, in one go (it ran at the start of the round: profiles/r02_first_call_summary.txt; the
# byte-form child process it also timed no longer exists — bench.py's e2e leg is in-process now — so those steps are gone):
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
#
# 1. the full -m gpu suite without -x (every shield removed), and the byte form under compute-sanitizer memcheck;
# 2. the default bench (e2e.packed8 = the byte form's end-to-end number, from its child process);
# 3. the persisting-L2 experiment of DESIGN.md §10 item 0: the same bench with the device limit raised;
# Everything lands in gpurun_out/r02_first/ — summarise into profiles/ afterwards (tools/ncu_summary.py).
set -u
OUT=gpurun_out/r02_first
mkdir -p "$OUT"

echo "== 1. the FULL -m gpu suite, no -x, no shields" | tee "$OUT/summary.txt"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/gpu_suite.log" 2>&1
echo "exit $?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/gpu_suite.log" | tee -a "$OUT/summary.txt"
grep -E "^(FAILED|ERROR)" "$OUT/gpu_suite.log" | tee -a "$OUT/summary.txt"

echo "== 1b. memcheck on one byte-form case + one mode-3 case" | tee -a "$OUT/summary.txt"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 \
  python -m pytest "tests/test_zz_packed8_gpu.py::test_byte_form_decodes_like_the_host_and_ticks_like_the_oracle[777-2-5]" \
  "tests/test_zz_packed8_gpu.py::test_tick_mode_3_consumes_the_bytes_itself[777-2-5]" \
  -m gpu -q -p no:cacheprovider > "$OUT/packed8_memcheck.log" 2>&1
echo "exit $?" | tee -a "$OUT/summary.txt"
grep -E "ERROR SUMMARY|passed|failed" "$OUT/packed8_memcheck.log" | tail -3 | tee -a "$OUT/summary.txt"

echo "== 2. default bench" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
echo "exit $?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_default.json" <<'EOF' | tee -a "$OUT/summary.txt"
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ticks/s", round(l["value"]), "us/step", round(1e3 * l["ms_per_step"], 2), "frac", round(l["roofline"]["frac"], 3))
    e = l["e2e"]
    print("e2e", round(e["value"]), "api", e["api"][:40], "| packed16", round(e["packed16"]["value"]), "| packed8", e.get("packed8"))
except Exception as ex:
    print("could not read the bench line:", ex)
EOF

echo "== 2c. the HBM-resident tick on byte frames (bench --inbox bytes, tick mode 3), kernels only" | tee -a "$OUT/summary.txt"
MRQ_BENCH_FAST=1 timeout 600 python bench.py --inbox bytes > "$OUT/bench_inbox_bytes.json" 2> "$OUT/bench_inbox_bytes.err"
echo "exit $?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_inbox_bytes.json" <<'EOF' | tee -a "$OUT/summary.txt"
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("bytes inbox: ticks/s", round(l["value"]), "us/step", round(1e3 * l["ms_per_step"], 2), "frac", round(l["roofline"]["frac"], 3),
          "|", l["config"]["inbox"])
except Exception as ex:
    print("could not read the bench line:", ex)
EOF

echo "== 3. persisting-L2 limit raised (MRQ_L2_PERSIST_MB=80), kernels only" | tee -a "$OUT/summary.txt"
for mb in 80 48; do
  MRQ_L2_PERSIST_MB=$mb MRQ_BENCH_FAST=1 timeout 600 python bench.py > "$OUT/bench_l2_$mb.json" 2> "$OUT/bench_l2_$mb.err"
  python - "$OUT/bench_l2_$mb.json" $mb <<'EOF' | tee -a "$OUT/summary.txt"
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("persist MB", sys.argv[2], "us/step", round(1e3 * l["ms_per_step"], 2), "frac", round(l["roofline"]["frac"], 3))
except Exception as ex:
    print("persist MB", sys.argv[2], "failed:", ex)
EOF
done
MRQ_BENCH_FAST=1 timeout 600 python bench.py --l2 0 > "$OUT/bench_l2_off.json" 2> "$OUT/bench_l2_off.err"

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv >> "$OUT/summary.txt" 2>&1
cat "$OUT/summary.txt"

echo "== 5. quorum-kernel shape explorer (tools/k3_explore.cu)" | tee -a "$OUT/summary.txt"
PEAK=$(python -c "import json;print(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])" 2>/dev/null || echo 6583.5)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/k3_explore tools/k3_explore.cu > "$OUT/k3_explore_build.log" 2>&1 \
  && timeout 300 /tmp/k3_explore "$PEAK" > "$OUT/k3_explore.txt" 2>&1
echo "exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/k3_explore.txt" 2>/dev/null | tee -a "$OUT/summary.txt"
