mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_planner.py -x -q -m gpu -s -k "nccl" 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
tail -c 1800 gpurun_out/r2h_bench_n2.json; tail -5 gpurun_out/r2h_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --candidates-total 8192 --no-extras > gpurun_out/r2h_bench_n2_strong.json 2> gpurun_out/r2h_bench_n2_strong.err
head -c 900 gpurun_out/r2h_bench_n2_strong.json; tail -3 gpurun_out/r2h_bench_n2_strong.err
