python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for k in mma g32h4; do
UDE_KERNEL=$k python bench.py --config c5 --series 200000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('c5 $k', '%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d['roofline']['kernel'][:60], 'loss', d['config']['loss_value'])" || tail -5 gpurun_out/b.err
done
