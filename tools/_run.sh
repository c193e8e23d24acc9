timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "search_graph" > gpurun_out/pytest86.log 2>&1; tail -15 gpurun_out/pytest86.log | cut -c1-250
