"""A/B of kernel families on plain NODE shapes that are not BASELINE configurations (tuning aid).
usage (under gpurun): python tools/node_shape_sweep.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_it(ude, torch, prob, theta, kernel, steps=10):
    os.environ["UDE_KERNEL"] = kernel
    try:
        dev = torch.device("cuda", 0)
        d_theta = torch.from_numpy(theta).to(dev)
        d_grad = torch.empty(prob.n_theta, dtype=torch.float64, device=dev)
        d_loss = torch.empty(1, dtype=torch.float64, device=dev)
        st = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(st)
        with ude.Engine(prob) as eng:
            kn = eng.kernel_name()
            for _ in range(3):
                eng.loss_grad_device(d_theta.data_ptr(), d_loss.data_ptr(), d_grad.data_ptr(), st.cuda_stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                eng.loss_grad_device(d_theta.data_ptr(), d_loss.data_ptr(), d_grad.data_ptr(), st.cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            return {"kernel": kn, "ms_per_eval": ms, "steps_per_s": eng.steps_per_eval() / (ms * 1e-3), "loss": float(d_loss.item()),
                    "grad_l1": float(d_grad.abs().sum().item())}
    finally:
        os.environ.pop("UDE_KERNEL", None)


def main():
    import torch
    import ude_b200 as ude
    P = ude.problem
    rng = np.random.default_rng(0)
    out = {}
    # d = 4 NODE without covariate, NN 4 -> 32 -> 4, multiple shooting: C3 minus the covariate
    Y, ts = ude.synthetic.lotka_volterra_data(rng, 100_000, 30, random_u0=True, pairs=2)
    p4 = P.Problem(d=4, nn_in=4, nn_hidden=32, nn_out=4, rhs_kind=P.RHS_NODE, solver=P.RK4, substeps=4, loss_kind=P.LOSS_MULTI_SHOOT,
                   pred_length=5).default_layout()
    ude.synthetic._tile_series(p4, Y, ts); p4.reg_weight = 1e-6; p4.finalize()
    th4 = ude.synthetic.theta_from(p4, ude.synthetic.init_pm(p4, rng), np.asarray(p4.data))
    out["d4_h32_multiple_shooting"] = {k: time_it(ude, torch, p4, th4, k) for k in ("node_d4_mma_h32", "node_d4_mma2_h32")}
    # d = 8 NODE, NN 8 -> 32 -> 8, 10^6 one-interval segments: C5 with a narrow hidden layer
    p8, th8 = ude.synthetic.config_c5(n_segments=1_000_000, hidden=32)
    out["d8_h32_one_interval_segments"] = {k: time_it(ude, torch, p8, th8, k) for k in ("node_d8_mma_h32", "node_d8_mma2_h32")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
