timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
