python bench.py > gpurun_out/b76.json 2>gpurun_out/b76.err; tail -c 300 gpurun_out/b76.err
python -c "
import json
d=json.loads(open('gpurun_out/b76.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['phases_ms_per_step'], d['single_query_latency_ms'], d['roofline']['frac'], d['map_index']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches76.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu76.log 2>&1; tail -c 200 gpurun_out/ncu76.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b76_ref.json 2>/dev/null; cut -c1-300 gpurun_out/b76_ref.json
