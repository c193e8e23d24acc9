#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter_window.py tests/test_gpu_filter_fast.py tests/test_gpu_tensorcore.py -x -q -m gpu > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_tc.log | cut -c1-300
for rows in 125000 1000000; do
timeout 300 python bench.py --rows $rows --steps 30 --warmup 5 --no-cpu-baseline --skip-map-index --skip-latency > gpurun_out/bench_shard_${rows}.json 2> gpurun_out/bench_shard.err || tail -c 400 gpurun_out/bench_shard.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_shard_${rows}.json").read().strip().splitlines()[-1])
print("rows $rows: %.0f QPS %.3f ms e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()})
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_125k.csv python bench.py --rows 125000 --steps 2 --warmup 3 --no-cpu-baseline --skip-latency --skip-map-index > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_125k.csv 2>/dev/null | grep -v "at::" | head -30
