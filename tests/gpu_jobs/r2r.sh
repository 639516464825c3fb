mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_9room.py -x -q -m gpu 2>&1 | tail -25
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_parity_9room.py 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('tree', d['value'], d['e2e']['value'], d['phase_ms_per_step'])"
