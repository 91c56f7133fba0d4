python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group_shapes_c3" 2>&1 | tail -15
for k in g8h4_ws_b2 mma_h32; do
UDE_KERNEL=$k python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('$k', '%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'loss', d['config']['loss_value'])" || tail -5 gpurun_out/b.err
done
