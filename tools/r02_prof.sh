#!/bin/bash
# ncu captures of every shipped hot kernel (one GPU; numbers printed under ncu are never bench values).
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/r02_prof.sh'
set -u
OUT=gpurun_out/r02_prof
mkdir -p "$OUT"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, skip, count, scenario args...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  timeout 600 $NCU -k "regex:$re" -s "$skip" -c "$cnt" -o "$OUT/$name" python tools/prof_r02.py "$@" > "$OUT/$name.log" 2>&1
  echo "$name exit $? $(ls -la $OUT/$name.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
cap prof_tick4_pertick_r02 tick_fast4 8 2 tick4          # launches 0.. : the 1st compacts; steady state from the 2nd on
cap prof_tick4_pertick_l2off_r02 tick_fast4 8 2 tick4 l2off
cap prof_tickslow4_r02 tick_slow4 8 1 tick4
cap prof_tick4_batched_r02 tick_fast4 1 1 tick4batch
cap prof_tick3_r02 tick_fast8 8 1 tick3
cap prof_tick0_r02 "tick_fast_kernel" 34 1 tick0        # (26 dry-run launches come first)
cap prof_k3_ldg256_r02 quorum_kernel_ldg256 1 1 k3 0
cap prof_k3_tma_r02 quorum_kernel_tma 1 1 k3 1
cap prof_k3_ldg128_r02 "^quorum_kernel_ldg$" 1 1 k3 2
# launch list of the default bench's kernels-only run
MRQ_BENCH_FAST=1 MRQ_BENCH_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file "$OUT/launches_r02.csv" python bench.py --steps 10 --warmup 3 > "$OUT/bench_under_ncu.log" 2>&1
echo "launch list exit $?"
MRQ_BENCH_FAST=1 MRQ_BENCH_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file "$OUT/launches_pertick_r02.csv" python bench.py --steps 10 --warmup 3 --graph off > "$OUT/bench_pertick_under_ncu.log" 2>&1
echo "per-tick launch list exit $?"
# the reports are ~28 MB each (gpurun_out/ carries 64 MiB back): summarise them HERE and keep only the summaries + raw CSV pages
mkdir -p "$OUT/md"
for f in "$OUT"/*.ncu-rep; do
  b=$(basename "$f" .ncu-rep)
  ncu -i "$f" --page raw --csv > "$OUT/$b.raw.csv" 2>/dev/null
  ncu -i "$f" --page source --csv 2>/dev/null | gzip > "$OUT/$b.source.csv.gz"
done
NCU_SUMMARY_OUT="$OUT/md" python tools/ncu_summary.py r02 "$OUT"/launches_r02.csv "$OUT"/*.ncu-rep > "$OUT/summary.log" 2>&1
mv "$OUT/md/r02_launches.md" "$OUT/md/r02_launches_batched.md" 2>/dev/null
NCU_SUMMARY_OUT="$OUT/md" python tools/ncu_summary.py r02 "$OUT"/launches_pertick_r02.csv >> "$OUT/summary.log" 2>&1
mv "$OUT/md/r02_launches.md" "$OUT/md/r02_launches_pertick.md" 2>/dev/null
rm -f "$OUT"/*.ncu-rep
ls -la "$OUT" "$OUT/md"
cat "$OUT"/md/*tick4_pertick_r02.md | head -60
