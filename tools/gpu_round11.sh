#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --rows 12500000 --dim 768 --queries 4096 --steps 5 --warmup 3 --bf16-only --skip-latency --skip-map-index --no-cpu-baseline > gpurun_out/bench_c5shard_bf16only.json 2> gpurun_out/bench_c5shard.err; echo "c5 shard rc=$?"
tail -c 600 gpurun_out/bench_c5shard.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_c5shard_bf16only.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("phases_ms_per_step"), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_filter_window.py -x -q -m gpu -k "exceptional or duplicates" > gpurun_out/sanitizer_memcheck_window.log 2>&1; echo "memcheck rc=$?"
tail -12 gpurun_out/sanitizer_memcheck_window.log | cut -c1-300
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
tail -6 gpurun_out/sanitizer_memcheck_smoke.log | cut -c1-300
