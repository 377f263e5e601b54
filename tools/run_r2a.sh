# round-2 evidence on one B200: tests, smoke, both bench arms, launch list (each ncu pass after the plain run exited 0)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench exit $?" >> gpurun_out/bench_r2a.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2a_ref.json 2> gpurun_out/bench_r2a_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|jb::" -c 400 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r2a.log 2>&1
