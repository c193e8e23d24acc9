python -m pytest tests -m gpu -x -q > gpurun_out/pytest90.log 2>&1; tail -3 gpurun_out/pytest90.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
python bench.py > gpurun_out/b90.json 2>gpurun_out/b90.err; tail -c 300 gpurun_out/b90.err
python -c "
import json
d=json.loads(open('gpurun_out/b90.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['phases_ms_per_step'], d['single_query_latency_ms']['cuda_graph'], d['roofline']['frac'], d['map_index']['value'], d['gpu_launches'], d['cpu_baseline']['value'])"
