mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2as_launches_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2as_bench_under_ncu.log 2>&1
