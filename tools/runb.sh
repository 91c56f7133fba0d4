python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 %.3e' % d['value']); print({k:(v.get('kernel'), '%.3e' % v.get('steps_per_s',0), round(v.get('fp64_frac',0),3)) for k,v in d['other_configs'].items()})"
