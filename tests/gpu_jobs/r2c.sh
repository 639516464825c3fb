set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_planner.py -x -q -m gpu -s -k "prune or pruned or chunked" 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d.get('value_pruned')); print(d.get('e2e_planner')); print(d.get('e2e_host_noise'))
PY
python bench.py --config seq --steps 5 --warmup 3 > gpurun_out/r2c_bench_seq.json 2> gpurun_out/r2c_bench_seq.err
tail -c 1500 gpurun_out/r2c_bench_seq.json; tail -3 gpurun_out/r2c_bench_seq.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_seq_launches.csv python bench.py --config seq --steps 1 --warmup 3 > gpurun_out/r2c_seq_under_ncu.log 2>&1
tail -3 gpurun_out/r2c_seq_under_ncu.log
