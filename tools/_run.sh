python -m pytest tests -m gpu -x -q > gpurun_out/pytest55.log 2>&1; tail -3 gpurun_out/pytest55.log
for rows in 125000 1000000; do
  python bench.py --rows $rows --steps 100 --no-cpu-baseline > gpurun_out/b55_${rows}.json 2>gpurun_out/b55.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/b55_${rows}.json').read().strip().splitlines()[-1]); print('rows', $rows, 'qps %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['phases_ms_per_step'])"
done
