mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tests/gpu_sanitize_probe.py 2>&1 | grep -E "ERROR SUMMARY|^ok|Invalid" | head -5
timeout 900 compute-sanitizer --tool racecheck python tests/gpu_sanitize_probe.py > gpurun_out/r2w_racecheck.txt 2>&1
grep -E "RACECHECK SUMMARY|^ok" gpurun_out/r2w_racecheck.txt
grep -E "Error: Race reported|Warning: Race reported|and (Read|Write) access" gpurun_out/r2w_racecheck.txt | sed -E 's/\+0x[0-9a-f]+//; s/\[[0-9]+ hazards\]//' | sort | uniq -c | sort -rn | head -12
