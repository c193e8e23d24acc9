#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for rows in 125000 250000 500000; do
for w in 0 1; do
HQ_FILTER_WINDOW=$w timeout 300 python bench.py --rows $rows --steps 30 --warmup 5 --no-cpu-baseline --skip-map-index --skip-latency > gpurun_out/bench_shard_${rows}_w$w.json 2> gpurun_out/bench_shard.err || tail -c 400 gpurun_out/bench_shard.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_shard_${rows}_w$w.json").read().strip().splitlines()[-1])
print("rows $rows window $w: %.0f QPS %.3f ms e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()})
PY
done
done
timeout 600 python -m pytest tests/test_gpu_filter_window.py tests/test_gpu_filter_fast.py -x -q -m gpu 2>&1 | tail -3
