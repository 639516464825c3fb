mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.256, .int.3" -s 18 -c 4 -o gpurun_out/r2al_lstm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2al_ncu.log 2>&1
ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.256, .int.0, .bool.1" -s 0 -c 64 -o gpurun_out/r2al_linear python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2al_ncu2.log 2>&1
ls -la gpurun_out/r2al*
