mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planner.py tests/test_gpu_parity_seq.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
python bench.py --no-cpu-baseline > gpurun_out/r2ae_bench_$i.json 2>/dev/null; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2ae_bench_$i.json') if l.startswith('{')][-1])
print(round(d['value']), round(d['e2e']['value']), round(d['value_pruned']['value']), d['e2e_planner']['ms_per_plan'], d['e2e_planner'].get('plan_ms'))
PY
done
