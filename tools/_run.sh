python -m pytest tests -m gpu -x -q > gpurun_out/pytest75.log 2>&1; tail -3 gpurun_out/pytest75.log | cut -c1-300
for i in 1 2; do python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1M', d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['clocks'])"; done
python bench.py --rows 125000 --steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('125k', d['value'], d['ms_per_step'], d['phases_ms_per_step'])"
