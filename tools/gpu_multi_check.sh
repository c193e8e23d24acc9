#!/usr/bin/env bash
# 2-GPU checks: cross-rank NCCL parity (tests/multi/nccl_check.py) + bench at 125 K and 500 K rows per GPU:  gpurun --gpus 2 -- bash tools/gpu_multi_check.sh
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/multi/nccl_check.py > gpurun_out/nccl_check_n2.txt 2>&1; echo "nccl_check rc=$?"
tail -6 gpurun_out/nccl_check_n2.txt | cut -c1-300
for rows in 250000 1000000; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --rows $rows --steps 40 --warmup 5 --no-cpu-baseline --skip-map-index > gpurun_out/bench_n2_$rows.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n2_$rows.json").read().strip().splitlines()[-1])
print("N=2 rows $rows: %.0f QPS %.3f ms e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()}, d["single_query_latency_ms"])
PY
done
