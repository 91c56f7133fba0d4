timeout 900 python -m pytest tests/test_gpu_host_api.py tests/test_gpu_ukf.py -m gpu -x -q 2>&1 | tail -8
