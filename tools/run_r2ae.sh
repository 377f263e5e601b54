# round-2 call ae (one B200): bridging pool of 2048 slots x 256 K trie nodes against 512 x 1 M (k_verdicts2 spent 6 % of its
# (build_p2048 = -DJB_POOL_SLOTS=2048 -DJB_TRIE_CAP_LOG2=18, the default since)
# samples asleep waiting for a slot and 7.5 % of its instructions polling, r2aa); clean and dirty workload, alternating
mkdir -p gpurun_out
run() {  # name, workload flag, env...
  name=$1; wl=$2; shift; shift
  env "$@" python bench.py --workload $wl --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2ae_$name.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2ae_$name.json')); t=d['detail']; print('$name', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'verdict', round(t['verdict_ms'],3), 'kwalk', round(t['kwalk_ms'],3), 'reruns', t.get('segment_reruns'))"
}
P=$PWD/jasper_b200/build_p2048/libjasper_b200_p2048.so
run base_1 clean JB_X=1
run p2048_1 clean JASPER_B200_LIB=$P
run base_2 clean JB_X=1
run p2048_2 clean JASPER_B200_LIB=$P
run dirty_base dirty JB_X=1
run dirty_p2048 dirty JASPER_B200_LIB=$P
JASPER_B200_LIB=$P python -m pytest tests -m gpu -x -q -k "bridge or reproduces or fuzz" > gpurun_out/r2ae_pytest_p2048.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ae_pytest_p2048.log
tail -2 gpurun_out/r2ae_pytest_p2048.log
