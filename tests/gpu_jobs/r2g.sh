mkdir -p gpurun_out
(cd tests/cuda && timeout 300 ./gemm_test) > gpurun_out/r2g_gemm_test_pair.txt 2>&1
(cd tests/cuda && GCPB200_NO_PAIR=1 timeout 300 ./gemm_test) > gpurun_out/r2g_gemm_test_nopair.txt 2>&1
grep -v "^T7" gpurun_out/r2g_gemm_test_pair.txt | tail -40
echo ---- nopair; grep "T6" gpurun_out/r2g_gemm_test_nopair.txt
