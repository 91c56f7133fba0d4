for k in mma_h32 t8b2 t8b3; do
UDE_KERNEL=$k python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('$k', '%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d['roofline']['kernel'][:60], d['config']['grad_l1'])" || tail -5 gpurun_out/b.err
done
