set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_planner.py -x -q -m gpu -s -k "prune or pruned or chunked" 2>&1 | tail -40 > gpurun_out/r2b_pruned_tests.log
cat gpurun_out/r2b_pruned_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 2500 gpurun_out/r2b_bench.json
tail -5 gpurun_out/r2b_bench.err
