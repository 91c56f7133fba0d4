python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print('f64 c3', '%.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])" || tail -5 gpurun_out/b.err
