python -m pytest tests -m gpu -x -q -k "golden or fuzz or mid_size or verdict" > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_s.log
python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/quick.json 2> gpurun_out/quick.err
JB_TRACE=1 python bench.py --no-cpu-baseline --steps 2 --warmup 2 > /dev/null 2> gpurun_out/trace7.err
