python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group_shapes" 2>&1 | tail -4
for k in "d2_g4h3" "d2_mma"; do
UDE_KERNEL=$k python bench.py --config c2 --series 100000 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('c2 $k', '%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'loss', d['config']['loss_value'])" || tail -5 gpurun_out/b.err
done
python bench.py --config c5 --series 200000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('c5', '%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'loss', d['config']['loss_value'])" || tail -5 gpurun_out/b.err
python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('c4', '%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'loss', d['config']['loss_value'])" || tail -5 gpurun_out/b.err
