mkdir -p gpurun_out
(cd tests/cuda && timeout 120 ./dec_tail3_test) > gpurun_out/r2o_dec_tail3_test.txt 2>&1; tail -12 gpurun_out/r2o_dec_tail3_test.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2o_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2o_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dec_tail3_kernel -s 8 -c 2 -o gpurun_out/r2o_dec_tail3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2o_ncu_tail.log 2>&1
tail -2 gpurun_out/r2o_ncu_tail.log; ls -la gpurun_out/r2o*
