# round-2 final evidence (after the pool and bridging changes) on one B200 (each ncu pass after the plain run exited 0): tests, smoke, both bench arms, launch list, ncu --set full of k_scan
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2ag_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ag_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/r2ag_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2ag_smoke.log
python bench.py > gpurun_out/r2ag_bench_n1.json 2> gpurun_out/r2ag_bench_n1.err; echo "bench exit $?" >> gpurun_out/r2ag_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ag_bench_n1_reference.json 2> gpurun_out/r2ag_bench_n1_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|jb::" -c 400 --csv --log-file gpurun_out/r2ag_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-dirty > gpurun_out/r2ag_ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:"k_scan" -c 1 -f -o /tmp/r2ag_scan python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-dirty > gpurun_out/r2ag_ncu_scan.log 2>&1
python tools/ncu_keymetrics.py /tmp/r2ag_scan.ncu-rep "r2ag: k_scan<1> (compact hash table), bench workload, ncu --set full --clock-control none" > gpurun_out/r2ag_scan_kernels_ncu_full.csv
