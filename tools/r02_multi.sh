#!/bin/bash
# One multi-GPU call (N = 2, 4 or 8 GPUs of one box): the sharded-engine parity tests, then bench.py with the three
# ways of handling the committed-index exchange (with and without the gather in the timed region).
#   /usr/local/graft/bin/gpurun --gpus N --timeout 1500 -- 'bash tools/r02_multi.sh N [notests]'
set -u
N=${1:-2}
OUT=gpurun_out/r02_multi_n$N
mkdir -p "$OUT"
if [ "${2:-}" != "notests" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_zz_compact_gpu.py -m gpu -q -p no:cacheprovider > "$OUT/multi_tests.log" 2>&1
  echo "multi tests exit $?"
  tail -4 "$OUT/multi_tests.log"
fi
PORT=29511
for G in ${GATHERS:-fused none nccl}; do
  PORT=$((PORT + 1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus "$N" --gather $G > "$OUT/bench_$G.json" 2> "$OUT/bench_$G.err"
  echo "bench --gather $G exit $?"
  python - "$OUT/bench_$G.json" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = l.get("e2e") or {}
    print("  ticks/s", round(l["value"]), "us/tick", round(l["ms_per_step"] * 1e3, 2), "frac", round(l["roofline"]["frac"], 3), "gather_check", l.get("gather_check"),
          "| e2e", round(e.get("value", 0)), "pack_us", round(e.get("pack_us_per_tick", 0)), "pre", round((e.get("preencoded") or {}).get("value", 0)))
except Exception as ex:
    print("  could not read the bench line:", ex)
PY
  tail -2 "$OUT/bench_$G.err"
done
if [ "${SKIP_PERTICK:-0}" = "1" ]; then exit 0; fi
# the same job, one launch pair per TICK (no batching): what the launch floor costs at this shard size
PORT=$((PORT + 1))
MRQ_BENCH_FAST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $PORT \
  bench.py --gpus "$N" --graph off > "$OUT/bench_pertick.json" 2> "$OUT/bench_pertick.err"
python - "$OUT/bench_pertick.json" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("per-tick launches: ticks/s", round(l["value"]), "us/tick", round(l["ms_per_step"] * 1e3, 2))
except Exception as ex:
    print("  could not read the per-tick bench line:", ex)
PY
