"""The two MultiCustomDerivatives examples of docs/src/MultipleTimeSeries.md on libude_cuda.

Example 1 (:27-61)  du_i/dt = r_i u_i NN(u_i): one growth rate PER SERIES              -> expression_rhs(..., series_param="r")
Example 2 (:63-92)  du/dt = NN([u; one_hot(i)]): the network is told which series it is -> one constant covariate per series
                                                                                           (one_hot_series_covariates), NODE trait
Both right-hand sides are compiled into kernels on first use; leave-future-out cross-validation of Example 1
trains all folds in one many-model handle.   usage (on a B200): python examples/multiple_time_series.py"""
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ude_b200 as ude  # noqa: E402


def logistic_series(m, n_pts, rng):
    """m populations with their own growth rates, logistic density dependence, multiplicative observation noise"""
    ts = np.linspace(0.0, 6.0, n_pts)
    rows = []
    for s in range(m):
        r, K, u = 0.4 + 0.25 * s, 2.0, 0.2 + 0.1 * rng.random()
        ys = []
        for k in range(n_pts):
            ys.append(u * np.exp(0.03 * rng.normal()))
            if k + 1 < n_pts:
                h = (ts[k + 1] - ts[k]) / 20
                for _ in range(20):
                    u = u + h * r * u * (1 - u / K)
        rows.append(pd.DataFrame({"time": ts, "series": s + 1, "x": ys}))
    return pd.concat(rows, ignore_index=True)


def main():
    rng = np.random.default_rng(1)
    # ---- Example 1: unique growth rates ----------------------------------------------------------------------------------------
    m = 2
    data = logistic_series(m, 40, rng)
    _, nn = ude.SimpleNeuralNetwork(1, 1, hidden=10)
    derivs = ude.expression_rhs(d=1, nn_out=1, series_param="r", du=["r*u1*y1"])                 # du .= p.r[index] .* u .* NN(u)
    model = ude.MultiCustomDerivatives(data, derivs, {"NN": nn, "r": np.full(m, 0.1)}, proc_weight=2.0, obs_weight=0.5, reg_weight=1e-4)
    t0 = time.perf_counter()
    ude.train_(model, loss_function="conditional likelihood", verbose=False, optim_options=dict(maxiter=800, step_size=0.02))
    t1 = time.perf_counter() - t0
    r_hat = model.process_model[-m:]
    print(f"example 1: {t1:.2f} s for 800 Adam iterations (incl. compiling the right-hand side); r_i * NN(K/2) per series:",
          np.round(r_hat * _nn_at(model, 1.0), 3), "(true r_i / 2 = [0.2, 0.325])")
    t0 = time.perf_counter()
    models, train, test, forecasts, trace = ude.leave_future_out_batched(model, 4, loss_function="conditional likelihood",
                                                                         optim_options=dict(maxiter=300, step_size=0.02))
    print(f"example 1: leave-future-out CV, 4 folds trained together in {time.perf_counter() - t0:.2f} s; final losses", np.round(trace[-1], 3))
    # ---- Example 2: one-hot series input ---------------------------------------------------------------------------------------
    m2 = 10
    Y, ts = ude.synthetic.lotka_volterra_data(rng, m2, 20, random_u0=True)
    data2 = pd.concat([pd.DataFrame({"time": ts, "series": s + 1, "x1": Y[s, 0], "x2": Y[s, 1]}) for s in range(m2)], ignore_index=True)
    X = ude.one_hot_series_covariates(data2)                                    # one_hot[index] = 1
    derivs2 = ude.expression_rhs(d=2, n_cov=m2, nn_out=2, du=["y1", "y2"])        # du .= NN(vcat(u, one_hot))
    _, nn2 = ude.SimpleNeuralNetwork(2 + m2, 2, hidden=10)
    model2 = ude.MultiCustomDerivatives(data2, derivs2, {"NN": nn2}, X=X, proc_weight=2.0, obs_weight=0.5, reg_weight=1e-4)
    t0 = time.perf_counter()
    ude.train_(model2, loss_function="conditional likelihood", verbose=False, optim_options=dict(maxiter=500, step_size=0.02))
    tr = model2.extra.get("loss_trace")
    print(f"example 2: {time.perf_counter() - t0:.2f} s for 500 Adam iterations with {m2} one-hot inputs; loss "
          + (f"{float(np.sum(tr[0])):.3f} -> {float(np.sum(tr[-1])):.3f}" if tr is not None else "finite"))
    assert np.all(np.isfinite(model.parameters)) and np.all(np.isfinite(model2.parameters))


def _nn_at(model, u):
    """the fitted network's output at state u (d = 1)"""
    p, pr = model.process_model, model.base_problem()
    H = pr.nn_hidden
    W1, b1 = p[pr.off_w1:pr.off_w1 + H], p[pr.off_b1:pr.off_b1 + H]
    W2, b2 = p[pr.off_w2:pr.off_w2 + H], p[pr.off_b2]
    return float(W2 @ np.tanh(W1 * u + b1) + b2)


if __name__ == "__main__":
    main()
