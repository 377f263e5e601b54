# round-2 call y (one B200): occupancy A/B of the verdict and walk kernels (launch bounds), alternating runs on one box
mkdir -p gpurun_out
for i in 1 2; do
  for v in base v1 v2 v3; do
    if [ $v = base ]; then unset JASPER_B200_LIB; else export JASPER_B200_LIB=$PWD/jasper_b200/build_$v/libjasper_b200_$v.so; fi
    python bench.py --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2y_${v}_$i.json 2>/dev/null
    python -c "import json; d=json.load(open('gpurun_out/r2y_${v}_$i.json')); print('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['detail']['verdict_ms'], d['detail']['kwalk_ms'])"
  done
done
for v in base v2 v3; do
  if [ $v = base ]; then unset JASPER_B200_LIB; else export JASPER_B200_LIB=$PWD/jasper_b200/build_$v/libjasper_b200_$v.so; fi
  python bench.py --workload dirty --no-dirty --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r2y_dirty_${v}.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2y_dirty_${v}.json')); print('dirty $v', d['ms_per_step'], d['e2e']['ms_per_step'], d['detail']['verdict_ms'], d['detail']['kwalk_ms'])"
done
