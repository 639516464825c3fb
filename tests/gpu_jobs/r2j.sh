mkdir -p gpurun_out
python bench.py --config seq --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('seq', d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('tree', d['value'], d['e2e']['value'], d['phase_ms_per_step'])"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 600 -c 7 -o gpurun_out/r2j_seq_gemm python bench.py --config seq --steps 1 --warmup 1 > gpurun_out/r2j_ncu.log 2>&1
tail -3 gpurun_out/r2j_ncu.log; ls -la gpurun_out/*.ncu-rep
