#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter_fast.py tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_filter.log | cut -c1-300
timeout 600 python tools/stress.py 11 60 > gpurun_out/stress_lazy.log 2>&1; echo "stress rc=$?"; tail -4 gpurun_out/stress_lazy.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("phases_ms_per_step"), d.get("roofline", {}).get("frac"), d["e2e"].get("result_gaps_ms"))
        print("   latency", d.get("single_query_latency_ms"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-latency --skip-map-index > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_r2.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki][:60]; v = float(r[vi].replace(",", ""))
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
    print("%-62s n=%3d total %.3f ms avg %.1f us" % (n, c, t / 1e6, t / c / 1e3))
PY
