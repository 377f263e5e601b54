ncu --set full --clock-control none --import-source on -k regex:k_walk -c 1 -f -o gpurun_out/walk_cur python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_walk_cur.log 2>&1
