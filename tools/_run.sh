python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard_ingest" 2>&1 | tail -2
python tools/db_build_time.py 1000000 1536
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches92.csv python tools/db_build_time.py 1000000 1536 > /dev/null 2>&1
grep -i "ingest" gpurun_out/launches92.csv | tail -1 | rev | cut -d, -f1 | rev
