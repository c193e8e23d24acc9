#!/usr/bin/env bash
# C2 (1 M x 1536, 1024-query batches) row-sharded over N GPUs of one box, then C5 on the same box:  gpurun --gpus N -- bash tools/scale_run.sh N
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --skip-map-index > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "c2 N=$N rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c2_n$N.json").read().strip().splitlines()[-1])
print("C2 N=$N: %.0f QPS %.3f ms e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()}, "per rank", [round(x, 3) for x in d["per_rank_ms_per_step"]], d["single_query_latency_ms"]["cuda_graph"], d["clocks"])
PY
bash tools/c5_run.sh $N
