mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_optim.py -x -q -m gpu -k "step_then" 2>&1 | tail -3
timeout 900 compute-sanitizer --tool racecheck python tests/gpu_sanitize_probe.py > gpurun_out/r2q_racecheck_full.txt 2>&1
grep -E "RACECHECK SUMMARY|^ok" gpurun_out/r2q_racecheck_full.txt
grep -E "Error: Race reported|and (Read|Write) access" gpurun_out/r2q_racecheck_full.txt | sed -E 's/\+0x[0-9a-f]+//; s/\[[0-9]+ hazards\]//' | sort | uniq -c | sort -rn | head -30
