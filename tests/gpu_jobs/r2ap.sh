mkdir -p gpurun_out
python -m pytest tests/test_gpu_planner.py -x -q -m gpu -s -k "nccl" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2ap_bench_n2.json 2> gpurun_out/r2ap_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ap_bench_n2.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','rank_consistent')}); print(d['e2e']['value'], d['e2e_host_noise']['value'], d['value_pruned']['value'], d['e2e_planner'])
PY
tail -3 gpurun_out/r2ap_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -c 700
