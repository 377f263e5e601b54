# A/B of k_scan build variants inside ONE gpurun call (boxes differ by a few %): tools/exp_variants.sh <tag> <lib>...
tag=$1; shift
for lib in "$@"; do
  n=$(basename $lib .so)
  JASPER_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/${tag}_$n.json 2> gpurun_out/${tag}_$n.err
done
