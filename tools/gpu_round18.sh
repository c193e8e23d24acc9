#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter_window.py tests/test_gpu_filter_fast.py -x -q -m gpu 2>&1 | tail -2
for v in 128 256; do
for rows in 1000000 125000; do
HQ_FILTER_WINDOW_CTA=$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v$v.csv python bench.py --rows $rows --steps 2 --warmup 3 --no-cpu-baseline --skip-latency --skip-map-index > gpurun_out/ncu_launch.log 2>&1; echo "variant $v rows $rows ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_v$v.csv 2>/dev/null | grep -E "cascade_win" 
done
done
