python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "--lanes 16 --hpl 2 --ws 1" "--lanes 8 --hpl 4 --ws 1" "--lanes 16 --hpl 2 --ws 2" "--lanes 8 --hpl 4 --ws 2"; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $cfg > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('$cfg', '%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4))" || tail -5 gpurun_out/b.err
done
