python -m pytest tests/test_gpu.py -m gpu -x -q -k "partitioned" > gpurun_out/pytest_part.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_part.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1v_n2.json 2> gpurun_out/bench_r1v_n2.err
