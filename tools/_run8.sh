n=8
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench85_n$n.json 2>gpurun_out/bench85_n$n.err
tail -c 300 gpurun_out/bench85_n$n.err
python -c "
import json
d=json.loads(open('gpurun_out/bench85_n$n.json').read().strip().splitlines()[-1]); print('N$n qps %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['phases_ms_per_step'], d['per_rank_ms_per_step'], d['single_query_latency_ms'])"
