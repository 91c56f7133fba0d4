python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in c4 c2; do
python bench.py --config $c --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$c value %.4e ms %.3f' % (d['value'], d['ms_per_step']), d['roofline']['kernel'][40:70])"
done
