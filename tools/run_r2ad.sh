# round-2 call ad (one B200): two-phase k_recount_probe (stretch list in shared memory, job decoding and walk out of line),
# (ran against the recount-inheritance experiment, commit 6b7a964, since reverted: JB_RECOUNT_NO_INHERIT / JB_RECOUNT_TEAM and the build_m3 / build_old variants exist only there)
# make_verdict's staging loads side by side; A/B against the committed build (build_old) on the same box
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "recount or fuzz or reproduces or verdict" > gpurun_out/r2ad_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ad_pytest_gpu.log
tail -2 gpurun_out/r2ad_pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2ad_$name.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2ad_$name.json')); t=d['detail']; print('$name', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'recount', round(t['recount_ms'],3), t['recount_probes'], 'kwalk', round(t['kwalk_ms'],3), 'verdict', round(t['verdict_ms'],3))"
}
M3=$PWD/jasper_b200/build_m3/libjasper_b200_m3.so
OLD=$PWD/jasper_b200/build_old/libjasper_b200_old.so
run old_1 JASPER_B200_LIB=$OLD
run new_1 JB_X=1
run new_m3_1 JASPER_B200_LIB=$M3
run new_plain JB_RECOUNT_NO_INHERIT=1
run new_t64 JB_RECOUNT_TEAM=64
run old_2 JASPER_B200_LIB=$OLD
run new_2 JB_X=1
run new_m3_2 JASPER_B200_LIB=$M3
