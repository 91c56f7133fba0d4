python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_api.py -m gpu -x -q -k "full_size or leave_future_out_predict" --durations=6 2>&1 | tail -12
