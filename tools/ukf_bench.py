"""Timing of the UKF marginal-likelihood loss+gradient (ude_ukf_loss_grad, host buffers in and out) on two shapes:
the C1 tutorial model (one series of 60 points: launch-bound) and a MultiNODE batch (many series: kernel-bound)."""
import json
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ude_b200 as ude  # noqa: E402


def run(name, prob, theta, reps):
    d = prob.d
    F = 0.3 * np.eye(d); Peta = 0.1 * np.eye(d)
    with ude.Engine(prob) as eng:
        for _ in range(3):
            L, g, gF = eng.ukf_loss_grad(theta, F, Peta, alpha=1.0)
        t0 = time.perf_counter()
        for _ in range(reps):
            L, g, gF = eng.ukf_loss_grad(theta, F, Peta, alpha=1.0)
        ms = (time.perf_counter() - t0) / reps * 1e3
    intervals = int(np.sum(np.maximum(prob.lengths - 1, 0)))
    steps = intervals * (2 * d + 1) * prob.substeps
    return {"case": name, "ms_per_eval": ms, "series": int(prob.n_series), "intervals": intervals, "sigma_points": 2 * d + 1,
            "rk_steps_per_eval": steps, "steps_per_s": steps / (ms * 1e-3), "loss": L}


out = []
p, th = ude.synthetic.config_c1()
out.append(run("C1 LV tutorial UDE, 1 series x 60 points", p, th, 50))
p, th = ude.synthetic.config_c2(n_series=10_000, n_pts=30)
out.append(run("NODE d=2 H=10, 10^4 series x 30 points", p, th, 10))
p, th = ude.synthetic.config_c3(n_series=2_000, n_pts=30)
out.append(run("MultiNODE d=4 +1 covariate H=32, 2000 series x 30 points", p, th, 10))
print(json.dumps(out, indent=1))
