#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for mq in 16 1; do
HQ_FILTER_WINDOW_MIN_Q=$mq timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_q1_$mq.csv python tools/latency_q1.py > gpurun_out/ncu_q1.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_q1.log | cut -c1-200
python - <<PY
import csv
rows = list(csv.reader(l for l in open("gpurun_out/launches_q1_$mq.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
names = [(r[ki][:64], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:]]
# the last search = the kernels after the last k_shard_ingest<12> that has Q=1 ... simply print the last 24 launches
tot = 0
for n, t in names[-26:]:
    print("   %-66s %8.1f us" % (n, t)); tot += t
print("min_q $mq: last 26 launches total %.1f us" % tot)
PY
done
