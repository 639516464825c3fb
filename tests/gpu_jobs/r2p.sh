mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_optim.py -x -q -m gpu -k "step_then" 2>&1 | tail -3
for tool in memcheck synccheck racecheck; do
  echo "## --tool $tool" >> gpurun_out/r2p_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python tests/gpu_sanitize_probe.py 2>&1 | grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|^ok" | head -20 >> gpurun_out/r2p_sanitizer.txt
done
cat gpurun_out/r2p_sanitizer.txt
