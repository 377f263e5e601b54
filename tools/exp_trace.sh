JB_TRACE=1 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/trace3.json 2> gpurun_out/trace3.err
