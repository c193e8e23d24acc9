timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench89_n2.json 2>gpurun_out/bench89_n2.err
tail -c 400 gpurun_out/bench89_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench89_n2.json').read().strip().splitlines()[-1]); print('N2 qps %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['phases_ms_per_step'], d['single_query_latency_ms']['cuda_graph'])"
