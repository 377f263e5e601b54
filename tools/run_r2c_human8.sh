# round-2 call c (8 x B200, one box): BASELINE configs[2] -- ONE 3.1 Gb assembly, k=37, p=4, sharded by chunk over 8 ranks
mkdir -p gpurun_out
{ echo "== nproc $(nproc)"; free -g | head -2; df -h /tmp /dev/shm; nvidia-smi --query-gpu=index,name,memory.total --format=csv; nvidia-smi topo -m | head -12; } > gpurun_out/r2c_box_probe_8gpu.log 2>&1
timeout 1000 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --config human --gpus 8 --steps 2 --warmup 1 --parity-mb 64 > gpurun_out/r2c_human3100_n8.json 2> gpurun_out/r2c_human3100_n8.err; echo "human exit $?" >> gpurun_out/r2c_human3100_n8.err
