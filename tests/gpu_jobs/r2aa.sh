mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2aa_bench.json') if l.startswith('{')][-1])
print(round(d['value']), d['phase_ms_per_step']); print(d['value_pruned']['value'], d['value_pruned']['ms_per_step'], d['value_pruned']['phase_ms_per_step'])
PY
