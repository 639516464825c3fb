mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tests/gpu_sanitize_probe.py 2>&1 | grep -E "ERROR SUMMARY|^ok|Invalid" | head -5
timeout 900 compute-sanitizer --tool racecheck python tests/gpu_sanitize_probe.py > gpurun_out/r2an_racecheck.txt 2>&1
grep -E "RACECHECK SUMMARY|^ok" gpurun_out/r2an_racecheck.txt
grep -E "Error: Race reported|Warning: Race reported|and (Read|Write) access" gpurun_out/r2an_racecheck.txt | sed -E 's/\+0x[0-9a-f]+//; s/\[[0-9]+ hazards\]//' | sort | uniq -c | sort -rn | head -12
timeout 600 compute-sanitizer --tool synccheck python tests/gpu_sanitize_probe.py 2>&1 | grep -E "ERROR SUMMARY|^ok" | head -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2an_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2an_bench_under_ncu.log 2>&1
PROBE_NCU=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2an_planner_launches_ncu.csv python tests/gpu_planner_probe.py > gpurun_out/r2an_probe_under_ncu.log 2>&1
python bench.py > gpurun_out/r2an_bench.json 2> gpurun_out/r2an_bench.err
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2an_bench_reference.json
python bench.py --config seq > gpurun_out/r2an_bench_seq.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2an_bench.json') if l.startswith('{')][-1])
print(round(d['value']), d['ms_per_step'], d['step_ms'], 'e2e', round(d['e2e']['value']), 'hn', round(d['e2e_host_noise']['value']), 'pruned', round(d['value_pruned']['value']), 'plan', d['e2e_planner']['plan_ms'], d['cpu_baseline'], d['roofline']['frac'], d['roofline_step'])
d=json.loads(open('gpurun_out/r2an_bench_reference.json').read()); print('ref', d['value'], d['cpu_baseline'])
d=json.loads([l for l in open('gpurun_out/r2an_bench_seq.json') if l.startswith('{')][-1]); print('seq', round(d['value']), d['ms_per_step'], d['roofline']['frac'], d.get('cpu_baseline'))
PY
