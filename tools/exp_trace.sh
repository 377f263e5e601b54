JB_TRACE=1 python bench.py --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/trace.json 2> gpurun_out/trace.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cur.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_cur.log 2>&1
