#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter_window.py tests/test_gpu_filter_fast.py tests/test_gpu_tensorcore.py -x -q -m gpu > gpurun_out/pytest_window.log 2>&1; echo "pytest window rc=$?"
tail -5 gpurun_out/pytest_window.log | cut -c1-400
timeout 600 python tools/stress.py 21 60 > gpurun_out/stress_win.log 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/stress_win.log | cut -c1-300
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
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_filter_cascade_win|k_filter_bits_tc|k_filter_predict" -s 12 -c 5 -o gpurun_out/r02_filter_win -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-latency --skip-map-index > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
