# round-2 call j (one B200): the scan index with the uniform minimizer ring and pair probes -- tests, A/B, ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "count_positions or scan_index or reproduces_reference_bytes or bench_contract" > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest_gpu.log
for i in 1 2; do
  python bench.py --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2j_ab_index_$i.json 2>gpurun_out/r2j_ab_index_$i.err
  JB_SCAN_NO_INDEX=1 python bench.py --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2j_ab_table_$i.json 2>/dev/null
done
ncu --set full --clock-control none -k regex:"k_scan" -c 1 -f -o /tmp/r2j_scan_index python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-dirty > gpurun_out/r2j_ncu_index.log 2>&1
python tools/ncu_keymetrics.py /tmp/r2j_scan_index.ncu-rep "r2j: k_scan through the scan index (k_scan<2>), bench workload, ncu --set full --clock-control none" > gpurun_out/r2j_scan_kernels_ncu_full.csv
