python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or full_size" 2>&1 | tail -3
for r in 1 0; do
UDE_HOST_CHUNK_EDGE=$r python bench.py --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('edge=$r value %.4e e2e %.4e ms %.3f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
