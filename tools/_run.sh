python -m pytest tests -m gpu -x -q > gpurun_out/pytest40.log 2>&1; tail -3 gpurun_out/pytest40.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/plain40.log 2>gpurun_out/plain40.err
python -c "
import json
d=json.loads(open('gpurun_out/plain40.log').read().strip().splitlines()[-1]); print(d['value'], d['phases_ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches40.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu40.log 2>&1
grep -E "k_filter_bits_tc|k_filter_cascade" gpurun_out/launches40.csv | tail -3 | awk -F'","' '{print $5, $NF}'
