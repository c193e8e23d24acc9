python -m pytest tests -m gpu -x -q > gpurun_out/pytest57.log 2>&1; tail -3 gpurun_out/pytest57.log
for rows in 125000 1000000; do
  python bench.py --rows $rows --steps 100 --no-cpu-baseline > gpurun_out/b57_${rows}.json 2>gpurun_out/b57.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/b57_${rows}.json').read().strip().splitlines()[-1]); print('rows', $rows, 'qps %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['phases_ms_per_step'])"
done
timeout 600 python bench_extra.py > gpurun_out/bench_extra57.jsonl 2>gpurun_out/bench_extra57.err; tail -c 300 gpurun_out/bench_extra57.err
cut -c1-600 gpurun_out/bench_extra57.jsonl
