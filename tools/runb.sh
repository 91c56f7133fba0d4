for k in g8h4_ws_b2 g8h4_ws_b3 g16h2_ws_b3; do
UDE_KERNEL=$k python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('$k', '%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4))" || tail -5 gpurun_out/b.err
done
