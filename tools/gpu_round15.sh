#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for mq in 16 1; do
HQ_FILTER_WINDOW_MIN_Q=$mq timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-map-index > gpurun_out/bench_lat_$mq.json 2> gpurun_out/bench_lat.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_lat_$mq.json").read().strip().splitlines()[-1])
print("min_q $mq:", d["single_query_latency_ms"])
PY
done
HQ_FILTER_WINDOW_MIN_Q=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_q1.csv python tools/latency_q1.py > gpurun_out/ncu_q1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_q1.log
python tools/launch_summary.py gpurun_out/launches_q1.csv 2>/dev/null | grep -v "at::" | head -30
