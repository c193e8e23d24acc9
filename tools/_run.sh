timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q > gpurun_out/pytest88.log 2>&1; tail -5 gpurun_out/pytest88.log | cut -c1-250
