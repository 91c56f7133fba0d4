python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "error_paths" 2>&1 | tail -12
