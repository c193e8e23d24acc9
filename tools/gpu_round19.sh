#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench_extra.py > gpurun_out/bench_extra.jsonl 2> gpurun_out/bench_extra.err; echo "bench_extra rc=$?"; tail -c 300 gpurun_out/bench_extra.err
python - <<'PY'
import json
for l in open("gpurun_out/bench_extra.jsonl"):
    try:
        d = json.loads(l)
        print(d.get("name"), {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ("value", "unit", "latency_ms_device", "ms_total", "ms_per_batch", "error", "bit_exact_round_trip", "bit_exact_inverse")}, d.get("roofline", {}).get("frac"), d.get("cuda_graph", {}).get("latency_ms_device"), d.get("clocks"))
    except Exception as e:
        print("bad line", e)
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-latency 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('quantized', d['map_index']['quantized']['ms_per_pass'], d['map_index']['quantized']['value'], 'map_index', d['map_index']['value'])"
