#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 500 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1_full.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("phases_ms_per_step"))
print("kernels", d.get("kernels_ms_per_step"))
print("roofline", {k: v for k, v in d["roofline"].items() if k not in ("note", "kernel", "timing")})
print("cpu_baseline", json.dumps(d.get("cpu_baseline"))[:600])
print("clocks", d.get("clocks"), "latency", d.get("single_query_latency_ms"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 300 gpurun_out/bench_ref.err; head -c 900 gpurun_out/bench_ref.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
