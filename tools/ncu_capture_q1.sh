#!/usr/bin/env bash
# Single-query path (one query against 1 M x 1536): launch list + one `ncu --set full` capture of its kernels.
#   gpurun -- bash tools/ncu_capture_q1.sh   -> gpurun_out/r02_launches_single_query.csv, gpurun_out/r02_q1.ncu-rep
set -u
mkdir -p gpurun_out
timeout 300 python tools/latency_q1.py > gpurun_out/q1_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_single_query.csv python tools/latency_q1.py > /dev/null 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_rerank_sparse_topk|k_filter_win_rows|k_filter_cascade_win|k_filter_predict" -s 8 -c 4 -o gpurun_out/r02_q1 -f python tools/latency_q1.py > gpurun_out/ncu_q1.log 2>&1; echo "q1 ncu rc=$?"
ls -la gpurun_out/r02_q1.ncu-rep
