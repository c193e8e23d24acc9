timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "search_graph" > gpurun_out/pytest80.log 2>&1; tail -30 gpurun_out/pytest80.log | cut -c1-250
python bench.py --steps 30 --no-cpu-baseline 2>gpurun_out/b80.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1M', d['value'], d['single_query_latency_ms'])"; tail -c 600 gpurun_out/b80.err
