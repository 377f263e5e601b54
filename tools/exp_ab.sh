for i in 1 2 3; do
  python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab_post_$i.json 2>/dev/null
  JB_NO_V2_POST=1 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab_nopost_$i.json 2>/dev/null
  JB_RECOUNT_ROWS=1 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab_rows_$i.json 2>/dev/null
done
