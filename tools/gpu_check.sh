#!/usr/bin/env bash
# Full GPU test suite + the default bench line (what the driver runs at round end) on one B200:  gpurun -- bash tools/gpu_check.sh
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("phases_ms_per_step"), d.get("roofline", {}).get("frac"), d["e2e"].get("result_gaps_ms"))
        print("   map_index", json.dumps(d.get("map_index"))[:1400])
        print("   latency", d.get("single_query_latency_ms"))
    except Exception as e:
        print(f, "unreadable", e)
PY
