ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches64_q1.csv python tools/latency_q1.py > gpurun_out/q1.log 2>&1
tail -3 gpurun_out/q1.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches64_q1.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-16:]:
    print(r[0], r[4][:80], r[-1])
PY
