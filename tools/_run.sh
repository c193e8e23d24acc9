python -m pytest tests -m gpu -x -q > gpurun_out/pytest43.log 2>&1; tail -3 gpurun_out/pytest43.log
for eh in 2; do
HQ_RERANK_EH=$eh python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/plain43_$eh.log 2>gpurun_out/plain43.err
python -c "
import json
d=json.loads(open('gpurun_out/plain43_$eh.log').read().strip().splitlines()[-1]); print('C2 eh=$eh', d['value'], d['phases_ms_per_step'])"
HQ_RERANK_EH=$eh python bench.py --rows 12500000 --dim 768 --queries 4096 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/c5shard43_$eh.json 2>gpurun_out/c5shard43.err
python -c "
import json
d=json.loads(open('gpurun_out/c5shard43_$eh.json').read().strip().splitlines()[-1]); print('C5 eh=$eh', d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['roofline']['frac'])"
done
