#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter_window.py tests/test_gpu_filter_fast.py -x -q -m gpu 2>&1 | tail -2
for mq in 16 1; do
HQ_FILTER_WINDOW_MIN_Q=$mq timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-map-index > gpurun_out/bench_lat_$mq.json 2> gpurun_out/bench_lat.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_lat_$mq.json").read().strip().splitlines()[-1])
print("min_q $mq: %.0f QPS %.3f ms" % (d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()}, d["single_query_latency_ms"])
PY
done
HQ_FILTER_WINDOW_MIN_Q=1 timeout 600 python -m pytest tests/test_gpu_filter_window.py tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-latency --skip-map-index > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_r2.csv 2>/dev/null | grep -E "filter|rerank_tc|predict" 
