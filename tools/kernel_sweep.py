"""A/B of compiled kernel specialisations on one resident workload (tuning aid; not a bench line).

usage: python tools/kernel_sweep.py [--config c3] [--series N] [--steps K] name[:ENV=VAL,...] ...
Each positional argument is a UDE_KERNEL substring, optionally followed by environment overrides for that run
(e.g. `mma2_h32_nt2:UDE_MAX_BLOCKS_PER_SM=1`).  Prints one JSON object per variant: steps/s (CUDA events around K resident
evaluations), the segment kernel's own time, loss and a gradient checksum (so two builds can be compared at full size).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--series", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--solver", default="rk4")
    ap.add_argument("--substeps", type=int, default=4)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("variants", nargs="+")
    a = ap.parse_args()
    import torch
    import bench
    import ude_b200 as ude
    args = argparse.Namespace(config=a.config, series=a.series, solver=a.solver, substeps=a.substeps)
    prob, theta0, name = bench.build_problem(ude, args, 0)
    dev = torch.device("cuda", 0)
    d_theta = torch.from_numpy(theta0).to(dev)
    d_grad = torch.empty(prob.n_theta, dtype=torch.float64, device=dev)
    d_loss = torch.empty(prob.n_models, dtype=torch.float64, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    peak = ude.measure_fp64_peak(0)
    print(json.dumps({"workload": name, "fp64_peak_tflops": peak, "flops_per_step": prob.flops_per_step()}), flush=True)
    for v in a.variants:
        kname, _, envs = v.partition(":")
        saved = {}
        over = {"UDE_KERNEL": kname}
        for kv in filter(None, envs.split(",")):
            k, _, val = kv.partition("=")
            over[k] = val
        for k, val in over.items():
            saved[k] = os.environ.get(k)
            os.environ[k] = val
        try:
            eng = ude.Engine(prob, device=0, dtype=1 if a.dtype == "f32" else 0)
            steps = eng.steps_per_eval()
            kn = eng.kernel_name()
            for _ in range(3):
                eng.loss_grad_device(d_theta.data_ptr(), d_loss.data_ptr(), d_grad.data_ptr(), ts.cuda_stream)
            torch.cuda.synchronize()
            eng.profile_begin(a.steps)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                eng.loss_grad_device(d_theta.data_ptr(), d_loss.data_ptr(), d_grad.data_ptr(), ts.cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            km = float(np.mean(eng.profile_read(a.steps)))
            rate = steps / (ms * 1e-3)
            out = {"variant": v, "kernel": kn, "ms_per_eval": ms, "kernel_ms": km, "steps_per_s": rate,
                   "fp64_frac": rate * prob.flops_per_step() / (peak * 1e12), "loss": float(d_loss.sum().item()),
                   "grad_l1": float(d_grad.abs().sum().item())}
            eng.close()
        except Exception as ex:      # a variant that does not exist must not stop the sweep
            out = {"variant": v, "error": repr(ex)}
        for k, val in saved.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
