mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_seq.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python bench.py --config seq --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('seq', d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"
