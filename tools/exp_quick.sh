python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_s.log
python bench.py --no-cpu-baseline > gpurun_out/quick.json 2> gpurun_out/quick.err
