nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/topo.txt; lscpu | grep -E "NUMA|Socket|^CPU\(s\)" >> gpurun_out/topo.txt
for a in 0 1; do
if [ $a = 1 ]; then export UDE_BENCH_NO_AFFINITY=1; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$a bench.py --gpus 8 --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/bench_n8_$a.json 2> gpurun_out/b8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n8_$a.json')); print('n8 noaff=$a', '%.4e steps/s' % d['value'], 'e2e %.4e' % d['e2e']['value'], d['config'].get('cpu_affinity'))" || tail -5 gpurun_out/b8.err
done
cat gpurun_out/topo.txt
