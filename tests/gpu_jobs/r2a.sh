set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_planner.py -x -q -m gpu -s 2>&1 | tail -40 > gpurun_out/r2a_planner_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
cat gpurun_out/r2a_planner_tests.log
