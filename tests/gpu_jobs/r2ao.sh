mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2ao_bench_n8.json 2> gpurun_out/r2ao_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ao_bench_n8.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','rank_consistent')}); print(d['e2e']['value'], d['e2e_host_noise']['value'], d['e2e_host_noise']['numa'], d['value_pruned']['value'], d['e2e_planner']['ms_per_plan'])
PY
tail -2 gpurun_out/r2ao_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --candidates-total 65536 --no-extras > gpurun_out/r2ao_bench_n8_strong.json 2> gpurun_out/r2ao_bench_n8_strong.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ao_bench_n8_strong.json') if l.startswith('{')][-1])
print('strong', {k:d.get(k) for k in ('value','ms_per_step','scaling','rank_consistent')}, d['e2e']['value'], d['config']['candidates_per_gpu'])
PY
tail -2 gpurun_out/r2ao_bench_n8_strong.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 5 --warmup 3 --no-extras > gpurun_out/r2ao_bench_n4.json 2> gpurun_out/r2ao_bench_n4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ao_bench_n4.json') if l.startswith('{')][-1])
print('n4', {k:d.get(k) for k in ('value','ms_per_step','rank_consistent')}, d['e2e']['value'])
PY
