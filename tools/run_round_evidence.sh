python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench_r1x.json 2> gpurun_out/bench_r1x.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1x_ref.json 2> gpurun_out/bench_r1x_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|jb::" -c 300 --csv --log-file gpurun_out/launches_r1x.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r1x.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_verdicts2" -c 5 -f -o gpurun_out/walk_r1x python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_walk_r1x.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_recount|k_events|k_verdicts$|k_concat|k_find" -c 9 -f -o gpurun_out/scan_r1x python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_scan_r1x.log 2>&1
