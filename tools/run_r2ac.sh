# round-2 call ac (one B200): the recount that inherits windows outside the edits, one warp per job -- tests that verify the
# (ran against the recount-inheritance experiment, commit 6b7a964, since reverted: JB_RECOUNT_NO_INHERIT / JB_RECOUNT_TEAM and the build_m3 / build_old variants exist only there)
# recount against a full scan, then an A/B of {inherit on/off} x {32, 64 threads per job} x {4, 3 CTAs per SM} on one box
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "recount or fuzz or reproduces" > gpurun_out/r2ac_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ac_pytest_gpu.log
tail -2 gpurun_out/r2ac_pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-dirty --steps 5 --warmup 3 > gpurun_out/r2ac_$name.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2ac_$name.json')); t=d['detail']; print('$name', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(t['recount_ms'],3), t['recount_probes'], round(t['kwalk_ms'],3))"
}
M3=$PWD/jasper_b200/build_m3/libjasper_b200_m3.so
run inherit_t32_1 JB_X=1
run plain_t64_1 JB_RECOUNT_NO_INHERIT=1 JB_RECOUNT_TEAM=64
run inherit_t64 JB_RECOUNT_TEAM=64
run plain_t32 JB_RECOUNT_NO_INHERIT=1
run inherit_t32_m3 JASPER_B200_LIB=$M3
run plain_t64_m3 JASPER_B200_LIB=$M3 JB_RECOUNT_NO_INHERIT=1 JB_RECOUNT_TEAM=64
run inherit_t32_2 JB_X=1
run plain_t64_2 JB_RECOUNT_NO_INHERIT=1 JB_RECOUNT_TEAM=64
