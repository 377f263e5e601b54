# round-2 call af (one B200): bridging search keeps a short path list in shared memory, make_verdict walks back by ballot;
# (build_old = the commit before, built by hand into jasper_b200/build_old/)
# A/B against the previous commit (build_old) on the same box, clean and dirty workload, then the GPU tests that reach those paths
mkdir -p gpurun_out
run() {  # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" python bench.py --workload $wl --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2af_$name.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2af_$name.json')); t=d['detail']; print('$name', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'verdict', round(t['verdict_ms'],3), 'kwalk', round(t['kwalk_ms'],3), 'reruns', t.get('segment_reruns'))"
}
OLD=$PWD/jasper_b200/build_old/libjasper_b200_old.so
run old_1 clean JASPER_B200_LIB=$OLD
run new_1 clean JB_X=1
run old_2 clean JASPER_B200_LIB=$OLD
run new_2 clean JB_X=1
run dirty_old dirty JASPER_B200_LIB=$OLD
run dirty_new dirty JB_X=1
python -m pytest tests -m gpu -x -q -k "bridge or reproduces or fuzz or verdict or golden" > gpurun_out/r2af_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2af_pytest_gpu.log
tail -2 gpurun_out/r2af_pytest_gpu.log
