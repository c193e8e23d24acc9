#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench_n1.err
timeout 900 python bench.py --rows 12500000 --dim 768 --queries 4096 --steps 5 --warmup 3 --bf16-only --skip-latency --no-cpu-baseline > gpurun_out/bench_c5shard_bf16only.json 2> gpurun_out/bench_c5shard.err; echo "c5 shard rc=$?"
tail -c 800 gpurun_out/bench_c5shard.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json", "gpurun_out/bench_c5shard_bf16only.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("phases_ms_per_step"), d.get("roofline", {}).get("frac"), d["e2e"].get("result_gaps_ms"))
        print("   map_index", json.dumps(d.get("map_index"))[:900])
    except Exception as e:
        print(f, "unreadable", e)
PY
