#!/usr/bin/env bash
# BASELINE config 5 on N GPUs of one box: 100 M x 768 rows row-sharded, 4096-query batches, bf16-only shards.
#   bash tools/c5_run.sh N [rows] [extra bench.py flags]
set -u
N=${1:-8}; ROWS=${2:-100000000}; shift; shift || true
mkdir -p gpurun_out
OUT=gpurun_out/c5_n${N}_rows${ROWS}
if [ "$N" -gt 1 ]; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N \
      bench.py --gpus $N --rows $ROWS --dim 768 --queries 4096 --steps 6 --warmup 3 --bf16-only --skip-latency --skip-map-index --no-cpu-baseline "$@" \
      > $OUT.json 2> $OUT.err
else
  timeout 1200 python bench.py --gpus 1 --rows $ROWS --dim 768 --queries 4096 --steps 6 --warmup 3 --bf16-only --skip-latency --skip-map-index --no-cpu-baseline "$@" \
      > $OUT.json 2> $OUT.err
fi
echo "c5 N=$N rc=$?"; tail -c 600 $OUT.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT.json").read().strip().splitlines()[-1])
    print("C5 N=$N rows=$ROWS: %.0f QPS, %.1f ms/step, e2e %.0f QPS" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["phases_ms_per_step"], "per rank", [round(x, 1) for x in d["per_rank_ms_per_step"]], "roofline frac", d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
