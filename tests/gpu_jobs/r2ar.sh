cp video_gcp_b200/libgcpb200.so /tmp/new.so
for rep in 1 2 3; do
for v in new prev; do
if [ $v = new ]; then cp /tmp/new.so video_gcp_b200/libgcpb200.so; else cp tests/cuda/libgcpb200_prev.so video_gcp_b200/libgcpb200.so; fi
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v', round(d['value']), d['phase_ms_per_step']['tree_recursion'], d['phase_ms_per_step']['rollout_total'], '| pruned', round(d['value_pruned']['value']), d['value_pruned']['phase_ms_per_step']['tree_recursion'], d['value_pruned']['phase_ms_per_step']['rollout_total'])"
done
done
cp /tmp/new.so video_gcp_b200/libgcpb200.so
