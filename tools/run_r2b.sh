# round-2 call b (one B200): the human-scale path at a quarter of the size, sanitizer logs on k37_mix, k_scan under ncu --set full
mkdir -p gpurun_out
timeout 900 python bench.py --config human --genome-mb 800 --steps 1 --warmup 1 --parity-mb 16 > gpurun_out/r2b_human800_n1.json 2> gpurun_out/r2b_human800_n1.err; echo "human exit $?" >> gpurun_out/r2b_human800_n1.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu.py -m gpu -x -q -k "test_polish_reproduces_reference_bytes and k37_mix" > gpurun_out/r2b_sanitizer_memcheck_k37_mix.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2b_sanitizer_memcheck_k37_mix.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu.py -m gpu -x -q -k "test_polish_reproduces_reference_bytes and k37_mix" > gpurun_out/r2b_sanitizer_racecheck_k37_mix.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2b_sanitizer_racecheck_k37_mix.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_pre_ncu.json 2>/dev/null && \
ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_recount" -c 3 -f -o gpurun_out/scan_r2b python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_scan_r2b.log 2>&1
python tools/ncu_keymetrics.py gpurun_out/scan_r2b.ncu-rep "r2b: compact table (8-byte slots), bench workload, ncu --set full --clock-control none" > gpurun_out/r2b_scan_kernels_ncu_full.csv
