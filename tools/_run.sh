python -m pytest tests -m gpu -x -q > gpurun_out/pytest50.log 2>&1; tail -6 gpurun_out/pytest50.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke50.log 2>&1; tail -2 gpurun_out/smoke50.log
python bench.py > gpurun_out/bench50.json 2>gpurun_out/bench50.err; tail -c 600 gpurun_out/bench50.err
python -c "
import json
d=json.loads(open('gpurun_out/bench50.json').read().strip().splitlines()[-1]); print('C2', d['value'], d['e2e'], d['phases_ms_per_step'], d['roofline']['frac'], d['map_index']['value'], d['cpu_baseline'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench50_ref.json 2>gpurun_out/bench50_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches50.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu50.log 2>&1
tail -3 gpurun_out/launches50.csv
