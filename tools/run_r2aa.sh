# round-2 call aa (one B200): source-level ncu capture of the pass-0 verdict and walk kernels (where do a walker's 38 k cycles per event go)
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dirty > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo "bench exit $?" >> gpurun_out/r2aa_bench.err
ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_verdicts" -c 3 -f -o gpurun_out/r2aa_walk python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-dirty > gpurun_out/r2aa_ncu.log 2>&1
ls -la gpurun_out/r2aa_walk.ncu-rep
