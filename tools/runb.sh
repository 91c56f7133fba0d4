python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'loss %.15g grad_l1 %.15g' % (d['config']['loss_value'], d['config']['grad_l1']))" || tail -5 gpurun_out/b.err
