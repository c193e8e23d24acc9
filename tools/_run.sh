python bench_extra.py > gpurun_out/extra81.jsonl 2>gpurun_out/extra81.err; tail -c 400 gpurun_out/extra81.err; cut -c1-420 gpurun_out/extra81.jsonl
