python -m pytest tests -m gpu -x -q > gpurun_out/pytest54.log 2>&1; tail -3 gpurun_out/pytest54.log
for rows in 125000 250000; do
 for th in 512 256; do
  HQ_LIST_CTA_THREADS=$th python bench.py --rows $rows --steps 50 --no-cpu-baseline > gpurun_out/b54_${rows}_${th}.json 2>gpurun_out/b54.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/b54_${rows}_${th}.json').read().strip().splitlines()[-1]); print('rows', $rows, 'threads', $th, 'qps %.0f ms %.3f' % (d['value'], d['ms_per_step']), d['phases_ms_per_step'])"
 done
done
python bench.py --steps 50 --no-cpu-baseline > gpurun_out/b54_1m.json 2>gpurun_out/b54.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/b54_1m.json').read().strip().splitlines()[-1]); print('1M qps %.0f ms %.3f' % (d['value'], d['ms_per_step']), d['phases_ms_per_step'])"
python bench_extra.py --only c1 2>&1 | tail -1 | cut -c1-400
