mkdir -p gpurun_out
(cd tests/cuda && timeout 300 ./gemm_test 2>&1 | tail -25)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2at_gpu_tests.log; cat gpurun_out/r2at_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2at_bench.json 2> gpurun_out/r2at_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2at_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print('e2e', d['e2e']['value'], 'host_noise', d['e2e_host_noise']['value'], 'pruned', d['value_pruned'], 'plan ms', d['e2e_planner']['ms_per_plan']); print(d.get('phases_ms'))
PY
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2at_bench.json') if l.startswith('{')][-1])
print(d['phase_ms_per_step']); print(d['value_pruned']['phase_ms_per_step'])
PY
python bench.py --config seq --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('seq', round(d['value']), d['ms_per_step'], d.get('recurrence_ms'), d['roofline']['frac'])"
