python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_api.py tests/test_gpu_ukf.py -m gpu -x -q -k "many_models or full_size or ensemble or batched or ukf or baseline_configs" 2>&1 | tail -3
python bench.py --config c4 --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c4 value %.4e ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']), d['roofline']['kernel'][:60])"
python bench.py --config c4 --solver tsit5 --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c4 tsit5 value %.4e ms %.3f' % (d['value'], d['ms_per_step']))"
