mkdir -p gpurun_out
for i in 1 2 3; do
python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2z_bench_$i.json 2> gpurun_out/r2z_bench_$i.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2z_bench_$i.json') if l.startswith('{')][-1])
print($i, round(d['value']), d['ms_per_step'], d['step_ms'], 'e2e', round(d['e2e']['value']), d['step_ms_e2e'], d['clocks'])
PY
done
python bench.py > gpurun_out/r2z_bench_full.json 2> gpurun_out/r2z_bench_full.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2z_bench_full.json') if l.startswith('{')][-1])
print('full', round(d['value']), d['ms_per_step'], d['step_ms'], 'e2e', round(d['e2e']['value']), d['step_ms_e2e'], d['clocks'], d['cpu_baseline'], d['value_pruned']['value'], d['e2e_planner']['ms_per_plan'])
PY
