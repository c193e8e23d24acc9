#!/usr/bin/env bash
# One `ncu --set full` capture of the step's kernels at HEAD (after a plain run of the same command exited 0) + the launch list:
#   gpurun -- bash tools/ncu_capture.sh        -> gpurun_out/r02_head_search.ncu-rep, gpurun_out/r02_head_item.ncu-rep, launches csv
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-latency"
timeout 600 $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_head.csv $CMD > /dev/null 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"k_rerank_tc|k_filter_bits_tc|k_filter_cascade_win|k_filter_predict" -s 20 -c 8 -o gpurun_out/r02_head_search -f $CMD --skip-map-index > gpurun_out/ncu_full.log 2>&1; echo "search ncu rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"k_item_pass_bulk|k_shard_ingest" -s 2 -c 3 -o gpurun_out/r02_head_item -f $CMD > gpurun_out/ncu_full2.log 2>&1; echo "item ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
