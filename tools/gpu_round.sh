#!/usr/bin/env bash
# One GPU-box call: the GPU suite (with the reference's own tests against the device classes), then the
# benchmark arms.  Everything worth keeping goes to gpurun_out/.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/host_cores.txt; free -g | head -2 >> gpurun_out/host_cores.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tests/conformance/run_reference_tests.py --out gpurun_out/conformance_all.json > gpurun_out/conformance_all.log 2>&1
tail -40 gpurun_out/conformance_all.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json", "gpurun_out/bench_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("phases_ms_per_step"), d.get("roofline", {}).get("frac"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "unreadable", e)
PY
