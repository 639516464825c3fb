mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_planner.py -x -q -m gpu -k "not 1024 and not nccl" 2>&1 | tail -12
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2s_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value']); print(d.get('value_pruned')); print(d.get('e2e_planner'))
PY
tail -3 gpurun_out/r2s_bench.err
