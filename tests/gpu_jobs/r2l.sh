mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_seq.py tests/test_gpu_parity.py tests/test_gpu_parity_train.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python tests/gpu_report.py 2>&1 | grep -E "e_df \(latents\)|images_df|== " | head -8
timeout 300 python bench.py --config seq --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('seq', d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('tree', d['value'], d['e2e']['value'], d['phase_ms_per_step'])"
