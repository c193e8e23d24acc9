python -m pytest tests -m gpu -x -q > gpurun_out/pytest46.log 2>&1; tail -12 gpurun_out/pytest46.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/plain46.log 2>gpurun_out/plain46.err
python -c "
import json
d=json.loads(open('gpurun_out/plain46.log').read().strip().splitlines()[-1]); print('C2', d['value'], d['phases_ms_per_step'], d['top1_hit_rate_perturbed'])"
python bench.py --rows 12500000 --dim 768 --queries 4096 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/c5shard46.json 2>gpurun_out/c5shard46.err
python -c "
import json
d=json.loads(open('gpurun_out/c5shard46.json').read().strip().splitlines()[-1]); print('C5', d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['roofline']['frac'])"
