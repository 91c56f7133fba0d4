nvidia-smi topo -m 2>&1 | head -8
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
for ex in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --exchange $ex --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/b2_$ex.json 2> gpurun_out/b2_$ex.err
python -c "
import json; d=json.load(open('gpurun_out/b2_$ex.json')); print('$ex', '%.4e steps/s' % d['value'], 'ms/step', round(d['ms_per_step'],4), d['gpu_launches'], d['config']['parallelism'], d['config']['loss_value'])" || tail -5 gpurun_out/b2_$ex.err
done
