"""A generated right-hand side against the hand-written catalog kernel it restates: same workload, CUDA events around resident
evaluations (tuning aid; not a bench line).  usage (under gpurun): python tools/user_rhs_bench.py [--series N] [--steps K]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--series", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    import torch
    import ude_b200 as ude
    P = ude.problem
    rng = np.random.default_rng(0)
    Y, ts = ude.synthetic.lotka_volterra_data(rng, a.series, 30, random_u0=True)
    t0 = time.perf_counter()
    user = ude.expression_rhs(d=2, nn_out=1, scalars=("r", "m", "theta"), du=["r*u1 - y1", "theta*y1 - m*u2"])
    ude.user_rhs.ensure_kernels(user, 10, 0)
    t_build = time.perf_counter() - t0
    out = {"workload": f"LV UDE (test/UDETests.jl:20-28), NN 2->10->1, multiple shooting pl=5, {a.series} series x 30 pts, RK4 x 4 substeps",
           "generate_and_compile_s": t_build}
    for label, kind, ur in (("catalog RhsLvUde", P.RHS_LV_UDE, None), ("generated from expressions", user.kind, user)):
        prob = P.Problem(d=2, nn_in=2, nn_hidden=10, nn_out=1, rhs_kind=kind, n_scalars=3, solver=P.RK4, substeps=4,
                         loss_kind=P.LOSS_MULTI_SHOOT, pred_length=5).default_layout()
        prob.user_rhs = ur
        ude.synthetic._tile_series(prob, Y, ts)
        prob.finalize()
        pm = ude.synthetic.init_pm(prob, np.random.default_rng(1), scalars=(1.0, 0.5, 0.5))
        theta = ude.synthetic.theta_from(prob, pm, np.asarray(prob.data))
        dev = torch.device("cuda", 0)
        d_theta = torch.from_numpy(theta).to(dev)
        d_grad = torch.empty(prob.n_theta, dtype=torch.float64, device=dev)
        d_loss = torch.empty(1, dtype=torch.float64, device=dev)
        st = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(st)
        with ude.Engine(prob) as eng:
            for _ in range(3):
                eng.loss_grad_device(d_theta.data_ptr(), d_loss.data_ptr(), d_grad.data_ptr(), st.cuda_stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                eng.loss_grad_device(d_theta.data_ptr(), d_loss.data_ptr(), d_grad.data_ptr(), st.cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            out[label] = {"kernel": eng.kernel_name(), "ms_per_eval": ms, "steps_per_s": eng.steps_per_eval() / (ms * 1e-3),
                          "loss": float(d_loss.item()), "grad_l1": float(d_grad.abs().sum().item())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
