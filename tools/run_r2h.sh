# round-2 call h (one B200): GPU tests, bench (both workloads), ncu --set full of the step's kernels (summary kept, report dropped)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest_gpu.log
python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "bench exit $?" >> gpurun_out/r2h_bench_n1.err
JB_TRACE=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-dirty > gpurun_out/r2h_trace.json 2> gpurun_out/r2h_trace.err
ncu --set full --clock-control none -k regex:"k_scan|k_recount|k_events|k_verdicts|k_walk|k_concat|k_find" -c 20 -f -o /tmp/r2h_step python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-dirty > gpurun_out/r2h_ncu_step.log 2>&1
python tools/ncu_keymetrics.py /tmp/r2h_step.ncu-rep "r2h: the kernels of one device step of the bench workload, ncu --set full --clock-control none" > gpurun_out/r2h_step_kernels_ncu_full.csv
