#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter_window.py -x -q -m gpu > gpurun_out/pytest_window.log 2>&1; echo "pytest window rc=$?"
tail -30 gpurun_out/pytest_window.log | cut -c1-400
timeout 300 python tools/filter_window_stats.py > gpurun_out/window_stats.txt 2>&1; cat gpurun_out/window_stats.txt | tail -12
timeout 300 python tools/filter_window_stats.py 12500000 768 512 > gpurun_out/window_stats_c5.txt 2>&1; cat gpurun_out/window_stats_c5.txt | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-map-index > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("phases_ms_per_step"), d.get("roofline", {}).get("frac"))
        print("   latency", d.get("single_query_latency_ms"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-latency --skip-map-index > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_r2.csv 2>/dev/null | grep -E "filter|rerank_tc|predict" 
