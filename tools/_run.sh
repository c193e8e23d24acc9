python -m pytest tests -m gpu -q > gpurun_out/pytest72.log 2>&1; tail -6 gpurun_out/pytest72.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke72.log 2>&1; tail -2 gpurun_out/smoke72.log | cut -c1-300
python bench.py > gpurun_out/b72.json 2>gpurun_out/b72.err; tail -c 300 gpurun_out/b72.err
python -c "
import json
d=json.loads(open('gpurun_out/b72.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['phases_ms_per_step'], d['cpu_baseline'])"
