python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('%.3e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],3), 'e2e %.3e' % d['e2e']['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), d['roofline']['kernel'], d['roofline']['traffic'])" || tail -5 gpurun_out/b.err
for K in 1 3 4 8 12; do UDE_HOST_CHUNKS=$K python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('K=$K e2e %.3e' % d['e2e']['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3))"; done
