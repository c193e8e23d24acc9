#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/multi/nccl_check.py > gpurun_out/nccl_check_n2.txt 2>&1; echo "nccl_check rc=$?"
tail -12 gpurun_out/nccl_check_n2.txt | cut -c1-400
bash tools/c5_run.sh 2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "quantize" 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-latency > gpurun_out/bench_n1b.json 2>gpurun_out/bench_n1b.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1b.json").read().strip().splitlines()[-1])
print("C2: %.0f QPS e2e %.0f" % (d["value"], d["e2e"]["value"]), d["e2e"]["result_gaps_ms"], "quantized", d["map_index"]["quantized"]["ms_per_pass"], d["map_index"]["quantized"]["value"])
PY
