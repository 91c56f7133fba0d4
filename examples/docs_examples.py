"""Two models of docs/src/examples.md on libude_cuda.

regime changes (:25-56)     derivs!: vals = NN([u1, u2, X1, t/50 - 1]); du = vals        -> time-dependent NODE, written here as expressions
                            with time_input = (1/50, -1) AND through the catalog entry node_time_rhs: same numbers
fishery (:150-185)          du1 = NN([u; X])[1]; du2 = r u2 (1 - u2/K) - q u1            -> expressions vs the catalog's FISHERY
then a leave-future-out ensemble (3 network seeds x 4 folds in one handle, the C4 configuration in miniature) and the
unscented-Kalman-filter fit (train!(...; loss_function = "marginal likelihood")).   usage (on a B200): python examples/docs_examples.py"""
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ude_b200 as ude  # noqa: E402


def main():
    # ---- regime changes: time-dependent NODE with a covariate --------------------------------------------------------------------
    data, X = ude.simulate_coral_data()
    data, hold = data.iloc[:80], data.iloc[79:90]
    _, nn = ude.SimpleNeuralNetwork(4, 2, hidden=10)
    derivs = ude.expression_rhs(d=2, n_cov=1, nn_out=2, time_input=(1.0 / 50.0, -1.0), du=["y1", "y2"])
    kw = dict(proc_weight=2.5, obs_weight=10.0, reg_weight=10 ** -2.5)
    model = ude.CustomDerivatives(data, derivs, {"NN": nn}, X=X, **kw)
    t0 = time.perf_counter()
    ude.gradient_descent_(model, maxiter=500)
    ude.gradient_descent_(model, maxiter=500, step_size=0.01)
    t_fit = time.perf_counter() - t0
    # the catalog entry of the same model runs on a kernel of another geometry (4 lanes x 3 hidden units instead of 2 x 5): the
    # summation order differs, so the two fits agree to round-off per evaluation, not bit for bit after 1000 Adam iterations
    a = ude.CustomDerivatives(data, derivs, {"NN": nn}, X=X, **kw)
    twin = ude.CustomDerivatives(data, ude.node_time_rhs(2, 1, 2, 1.0 / 50.0, -1.0), {"NN": nn}, X=X, **kw)
    ude.gradient_descent_(a, maxiter=20); ude.gradient_descent_(twin, maxiter=20)
    same = float(np.max(np.abs(a.parameters - twin.parameters)) / np.max(np.abs(twin.parameters)))
    inits, obs, preds = ude.predictions(model, hold)
    print(f"regime changes: 2 x 500 Adam iterations in {t_fit:.2f} s; generated vs catalog kernels after 20 iterations: {same:.1e}; "
          f"one-step RMSE on 10 held-out points {float(np.sqrt(np.mean((preds - obs) ** 2))):.3f}")
    # ---- fishery: parametric logistic growth + network harvest --------------------------------------------------------------------
    rng = np.random.default_rng(4)
    ts = np.arange(1960.0, 2000.0)
    Xv = (ts >= 1992).astype(float)
    H, y = np.zeros(40), np.zeros(40)
    h, b = 0.3, 3.0
    for k in range(40):                     # harvest relaxes to an entry-dependent level; logistic stock minus harvest (20 Euler substeps)
        H[k], y[k] = h * np.exp(0.03 * rng.normal()), b * np.exp(0.03 * rng.normal())
        for _ in range(20):
            h, b = h + 0.05 * 0.5 * ((0.45 - 0.25 * Xv[k]) - h), b + 0.05 * (0.5 * b * (1 - b / 4.0) - 1.0 * h)
    state = pd.DataFrame({"t": ts, "H": H, "y": y}); cov = pd.DataFrame({"t": ts, "X": Xv})
    _, nn3 = ude.SimpleNeuralNetwork(3, 1, hidden=10)
    fish = ude.expression_rhs(d=2, n_cov=1, nn_out=1, scalars=("q", "r", "K"), du=["y1", "r*u2*(1 - u2/K) - q*u1"])
    init = {"NN": nn3, "q": 1.0, "r": 0.5, "K": 4.0}
    kw = dict(time_column_name="t", proc_weight=2.0, obs_weight=0.5, reg_weight=1e-4)
    m_f = ude.CustomDerivatives(state, fish, init, X=cov, **kw)
    ude.gradient_descent_(m_f)
    m_c = ude.CustomDerivatives(state, ude.FISHERY, init, X=cov, **kw)
    ude.gradient_descent_(m_c)
    same2 = float(np.max(np.abs(m_f.parameters - m_c.parameters)) / np.max(np.abs(m_c.parameters)))
    print(f"fishery: generated vs catalog kernels after 500 iterations: {same2:.1e}; q, r, K =", np.round(m_f.process_model[-3:], 3),
          "(simulated with 1.0, 0.5, 4.0)")
    # ---- ensemble x folds in one handle, then the Kalman-filter fit -----------------------------------------------------------------
    t0 = time.perf_counter()
    models, train, test, fc, trace = ude.leave_future_out_batched(m_c, 4, loss_function="conditional likelihood", member_seeds=[1, 2, 3],
                                                                  optim_options=dict(maxiter=300, step_size=0.05))
    print(f"fishery: 3 network seeds x 4 leave-future-out folds = {len(models)} models trained together, 300 iterations each, "
          f"in {time.perf_counter() - t0:.2f} s; mean final loss {float(np.mean(trace[-1])):.3f}")
    t0 = time.perf_counter()
    Pnu, Px = ude.train_(m_c, "marginal likelihood", verbose=False, loss_options=dict(process_error=0.1, observation_error=0.05),
                         optim_options=dict(maxiter=200, step_size=0.02))
    tr = m_c.extra["loss_trace"]
    print(f"fishery: marginal likelihood (UKF), 200 Adam iterations in {time.perf_counter() - t0:.2f} s; -loglik {tr[0]:.2f} -> {tr[-1]:.2f}; "
          f"P_nu diag {np.round(np.diag(Pnu), 4)}")
    assert same < 1e-9 and same2 < 1e-6 and np.all(np.isfinite(tr))


if __name__ == "__main__":
    main()
