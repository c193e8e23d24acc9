python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard_ingest" > gpurun_out/pytest78.log 2>&1; tail -25 gpurun_out/pytest78.log | cut -c1-250
python tools/db_build_time.py 1000000 1536
python tools/db_build_time.py 12500000 768
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches78.csv python tools/db_build_time.py 1000000 1536 > /dev/null 2>&1
grep -i "ingest" gpurun_out/launches78.csv | tail -2 | cut -c1-50,250-400
