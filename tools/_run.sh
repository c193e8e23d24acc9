timeout 600 python tools/stress.py 0 40 > gpurun_out/stress69a.log 2>&1; grep -c "^ok" gpurun_out/stress69a.log; grep -v "^ok" gpurun_out/stress69a.log | cut -c1-260
timeout 600 python tools/stress.py 7 60 > gpurun_out/stress69b.log 2>&1; grep -c "^ok" gpurun_out/stress69b.log; grep -v "^ok" gpurun_out/stress69b.log | cut -c1-260
