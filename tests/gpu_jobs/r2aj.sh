mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.256, .int.3" -s 18 -c 4 -o gpurun_out/r2aj_lstm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2aj_ncu.log 2>&1
tail -2 gpurun_out/r2aj_ncu.log; ls -la gpurun_out/r2aj*
