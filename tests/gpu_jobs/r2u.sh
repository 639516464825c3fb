mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 95 -c 12 -o gpurun_out/r2u_tree_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2u_ncu.log 2>&1
tail -2 gpurun_out/r2u_ncu.log
python bench.py --steps 3 --warmup 3 --candidates-total 65536 --no-extras --no-cpu-baseline 2>/dev/null > gpurun_out/r2u_bench_n1_strong.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2u_bench_n1_strong.json') if l.startswith('{')][-1])
print('strong N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['rollout_chunk'])
PY
