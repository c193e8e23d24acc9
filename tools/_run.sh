python -m pytest tests -m gpu -x -q > gpurun_out/pytest84.log 2>&1; tail -3 gpurun_out/pytest84.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
python bench.py > gpurun_out/b84.json 2>gpurun_out/b84.err; tail -c 300 gpurun_out/b84.err
python -c "
import json
d=json.loads(open('gpurun_out/b84.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['phases_ms_per_step'], d['single_query_latency_ms'], d['roofline']['frac'], d['map_index']['value'], d['gpu_launches'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches84.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu84.log 2>&1
