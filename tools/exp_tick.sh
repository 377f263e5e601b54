python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_s.log
for t in 0 15 16 17 18; do
  JB_WALK_TICK=$t python bench.py --no-cpu-baseline > gpurun_out/tick_$t.json 2> gpurun_out/tick_$t.err
done
