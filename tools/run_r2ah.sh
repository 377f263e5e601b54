# round-2 call ah (one B200): k_events2 (one mask word per lane from counts staged in shared memory) against k_events
# (k_events2 and JB_EVENTS_V1 existed only in commit eac673c, removed in 1dd74e7: see profiles/r2ah_events_ab.txt)
# (JB_EVENTS_V1); the full GPU test suite and smoke() run on the build with k_events2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2ah_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ah_pytest_gpu.log
tail -2 gpurun_out/r2ah_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/r2ah_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2ah_smoke.log
tail -2 gpurun_out/r2ah_smoke.log
run() {  # name, env...
  name=$1; shift
  env "$@" JB_TRACE_KERNELS=1 python bench.py --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2ah_$name.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2ah_$name.json')); t=d['detail']; print('$name', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'scan', round(t['scan_ms'],3), 'recount', round(t['recount_ms'],3), 'verdict', round(t['verdict_ms'],3), 'kwalk', round(t['kwalk_ms'],3))"
}
run v1_1 JB_EVENTS_V1=1
run v2_1 JB_X=1
run v1_2 JB_EVENTS_V1=1
run v2_2 JB_X=1
