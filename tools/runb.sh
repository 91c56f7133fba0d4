SECONDS=0
python bench.py 2> gpurun_out/b.err > gpurun_out/bench_default.json; echo "bench wall seconds: $SECONDS"; tail -3 gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print('%.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['roofline']['frac'], d['clocks']); print(json.dumps(d['other_configs'], indent=0)[:2500])"
