mkdir -p gpurun_out
BENCH_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
grep e2e gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2e_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e'])
PY
