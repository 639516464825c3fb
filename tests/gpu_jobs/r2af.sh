BENCH_TRACE=1 python bench.py --no-cpu-baseline 2>&1 >/dev/null | grep "^\[plan\]"
