python -m pytest tests/test_gpu_host_api.py -m gpu -x -q 2>&1 | tail -6
