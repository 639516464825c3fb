mkdir -p gpurun_out
python tests/gpu_cluster_probe.py 2>&1 | tee gpurun_out/r2f_cluster_probe.txt
