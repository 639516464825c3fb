mkdir -p gpurun_out
python tests/gpu_planner_probe.py > gpurun_out/r2d_probe.txt 2>&1
cat gpurun_out/r2d_probe.txt | cut -c1-200
