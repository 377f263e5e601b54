# round-2 call g (one B200): k_scan with tiles by TMA bulk copy against per-lane staging, A/B on one box; ncu --set full of the step's kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest_gpu.log
for i in 1 2 3; do
  python bench.py --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2g_ab_tma_$i.json 2>/dev/null
  JASPER_B200_LIB=$PWD/jasper_b200/build_notma/libjasper_b200_notma.so python bench.py --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2g_ab_notma_$i.json 2>/dev/null
done
ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_recount|k_events|k_verdicts|k_walk|k_concat|k_place|k_seg" -c 24 -f -o gpurun_out/r2g_step python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-dirty > gpurun_out/r2g_ncu_step.log 2>&1
python tools/ncu_keymetrics.py gpurun_out/r2g_step.ncu-rep "r2g: every kernel of one step (device step of the bench workload), ncu --set full --clock-control none" > gpurun_out/r2g_step_kernels_ncu_full.csv
