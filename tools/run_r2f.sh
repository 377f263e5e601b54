# round-2 call e (one B200): GPU tests, bench, host timeline of one step, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest_gpu.log
python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench exit $?" >> gpurun_out/r2f_bench_n1.err
JB_TRACE=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_trace.json 2> gpurun_out/r2f_trace.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|jb::" -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_ncu_launches.log 2>&1
