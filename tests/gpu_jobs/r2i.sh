mkdir -p gpurun_out
python tests/gpu_report.py > gpurun_out/r2i_gpu_report.txt 2>&1; tail -25 gpurun_out/r2i_gpu_report.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2i_gpu_tests.log; cat gpurun_out/r2i_gpu_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['phase_ms_per_step'])"
