ncu --set full --clock-control none --import-source on -k regex:"k_filter_bits_tc|k_filter_cascade_lists|k_rerank_tc" -s 12 -c 4 -o /tmp/prof_head python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu95.log 2>&1
python tools/ncu_summary.py /tmp/prof_head.ncu-rep > gpurun_out/ncu95_summary.txt 2>gpurun_out/ncu95_summary.err
grep -E "^\[|gpu__time_duration|issue_active|dram_throughput|tensor_cycles_active|inst_executed" gpurun_out/ncu95_summary.txt | cut -c1-130
