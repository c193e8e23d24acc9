python bench.py > gpurun_out/bench66.json 2>gpurun_out/bench66.err; tail -c 300 gpurun_out/bench66.err
python -c "
import json
d=json.loads(open('gpurun_out/bench66.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['map_index']['value'], d['single_query_latency_ms'], d['roofline']['frac'], d['cpu_baseline']['value'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench66_ref.json 2>gpurun_out/bench66_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches66.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu66.log 2>&1
timeout 600 python bench_extra.py > gpurun_out/bench_extra66.jsonl 2>gpurun_out/bench_extra66.err
cut -c1-260 gpurun_out/bench_extra66.jsonl
