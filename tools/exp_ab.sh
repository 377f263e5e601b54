python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_s.log
for i in 1 2 3; do
  python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab_a_$i.json 2>/dev/null
done
