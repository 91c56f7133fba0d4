timeout 600 python tools/ukf_bench.py > gpurun_out/ukf_r01.json 2> gpurun_out/ukf.err; tail -3 gpurun_out/ukf.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
