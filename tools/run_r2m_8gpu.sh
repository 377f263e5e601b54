# round-2 call m (8 x B200, one box): configs[2] with the final build, configs[3] (HiFi reads, p=2), configs[4] (one individual's database per GPU, 800 Mb reference)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29521 bench.py --config human --gpus 8 --steps 2 --warmup 1 --parity-mb 64 > gpurun_out/r2m_human3100_n8.json 2> gpurun_out/r2m_human3100_n8.err; echo "human exit $?" >> gpurun_out/r2m_human3100_n8.err
timeout 600 $T --master-port 29522 bench.py --config human --reads hifi --gpus 8 --steps 2 --warmup 1 --parity-mb 32 > gpurun_out/r2m_hifi3100_n8.json 2> gpurun_out/r2m_hifi3100_n8.err; echo "hifi exit $?" >> gpurun_out/r2m_hifi3100_n8.err
timeout 600 $T --master-port 29523 bench.py --config personalized --genome-mb 800 --gpus 8 --steps 2 --warmup 1 --parity-mb 16 > gpurun_out/r2m_personalized800_n8.json 2> gpurun_out/r2m_personalized800_n8.err; echo "personalized exit $?" >> gpurun_out/r2m_personalized800_n8.err
