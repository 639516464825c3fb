mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tests/gpu_sanitize_probe.py 2>&1 | grep -E "ERROR SUMMARY|^ok|Invalid" | head -5
timeout 900 compute-sanitizer --tool racecheck python tests/gpu_sanitize_probe.py > gpurun_out/r2ab_racecheck.txt 2>&1
grep -E "RACECHECK SUMMARY|^ok" gpurun_out/r2ab_racecheck.txt
grep -E "Error: Race reported|Warning: Race reported|and (Read|Write) access" gpurun_out/r2ab_racecheck.txt | sed -E 's/\+0x[0-9a-f]+//; s/\[[0-9]+ hazards\]//' | sort | uniq -c | sort -rn | head -12
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2ab_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2ab_bench_under_ncu.log 2>&1
PROBE_NCU=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2ab_planner_launches_ncu.csv python tests/gpu_planner_probe.py > gpurun_out/r2ab_probe_under_ncu.log 2>&1
ls -la gpurun_out/r2ab*
