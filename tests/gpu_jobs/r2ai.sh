for v in 0 1; do
if [ $v = 1 ]; then export EXP_NOHID=1; fi
python bench.py --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v', round(d['value']), d['phase_ms_per_step'])"
done
