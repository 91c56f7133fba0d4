UDE_PEER_TIMEOUT_MS=15000 timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q -k "threads or peer or two_or_more" 2>&1 | tail -8
