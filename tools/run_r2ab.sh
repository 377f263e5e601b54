# round-2 call ab (one B200): recount that inherits the windows outside the edits + walker with job slots in blocks; tests, A/B, full bench with parity
# (ran against the recount-inheritance experiment, commit 6b7a964, since reverted: JB_RECOUNT_NO_INHERIT / JB_RECOUNT_TEAM and the build_m3 / build_old variants exist only there)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ab_pytest_gpu.log
tail -3 gpurun_out/r2ab_pytest_gpu.log
for i in 1 2; do
  for v in new noinherit; do
    if [ $v = new ]; then unset JB_RECOUNT_NO_INHERIT; else export JB_RECOUNT_NO_INHERIT=1; fi
    python bench.py --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2ab_${v}_$i.json 2>/dev/null
    python -c "import json; d=json.load(open('gpurun_out/r2ab_${v}_$i.json')); t=d['detail']; print('$v', d['ms_per_step'], d['e2e']['ms_per_step'], t['recount_ms'], t['recount_probes'], t['kwalk_ms'], t['verdict_ms'])"
  done
done
unset JB_RECOUNT_NO_INHERIT
python bench.py > gpurun_out/r2ab_bench_n1.json 2> gpurun_out/r2ab_bench_n1.err; echo "bench exit $?" >> gpurun_out/r2ab_bench_n1.err
python -c "import json; d=json.load(open('gpurun_out/r2ab_bench_n1.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['cpu_baseline']['parity'][:80], '| dirty', d['dirty_workload']['ms_per_step'], d['dirty_workload']['cpu_baseline']['parity'][:60])"
