timeout 500 python tools/stress.py 0 40 > gpurun_out/stress68.log 2>&1; tail -45 gpurun_out/stress68.log | cut -c1-220
