python -m pytest tests -m gpu -x -q > gpurun_out/pytest93.log 2>&1; tail -3 gpurun_out/pytest93.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 50 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1M', d['value'], d['e2e']['value'], d['phases_ms_per_step'])"
