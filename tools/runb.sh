timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 --no-other-configs > gpurun_out/bench_n8.json 2> gpurun_out/b8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json')); print('n8', '%.4e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],4), d['gpu_launches'], d['config']['parallelism'], 'e2e %.4e' % d['e2e']['value'], d['clocks'])" || tail -5 gpurun_out/b8.err
