"""Expression right-hand sides used by tests/test_user_rhs.py (CPU) and tests/test_gpu_user_rhs.py (GPU)."""
import numpy as np

import cases


def lv_twin(ude):
    """the catalog's LV UDE (test/UDETests.jl:20-28; csrc RhsLvUde<0>) written as expressions"""
    return ude.expression_rhs(d=2, nn_out=1, scalars=("r", "m", "theta"), du=["r*u1 - y1", "theta*y1 - m*u2"], name="LV_TWIN")


def epidemic(ude):
    """not in the catalog: three states, one covariate, a time input, two network outputs, nonlinear in states, parameters AND
    network outputs (so the VJP needs the stashed y), with abs / exp / a rational term"""
    return ude.expression_rhs(
        d=3, n_cov=1, nn_out=2, scalars=("beta", "gamma", "K"), time_input=(0.5, -0.25),
        du=["-beta*u1*u2/(1 + u1**2) + y1*x1",
            "beta*u1*u2/(1 + u1**2) - gamma*u2*abs(y2) + 0.1*exp(-u3)",
            "gamma*u2*abs(y2) - u3*y1**2/K"], name="EPIDEMIC")


def growth_twin(ude):
    """the catalog's per-series growth model (docs/src/MultipleTimeSeries.md:51-55; csrc RhsSeriesGrowth<1>) as an expression
    with one trainable parameter per series"""
    return ude.expression_rhs(d=1, nn_out=1, series_param="r", du=["r*u1*y1"], name="GROWTH_TWIN")


def register(ude, name, rhs, sc0):
    """make the right-hand side reachable through cases.make_case(name, ...)"""
    cases.RHS_SPECS[name] = (rhs.kind, rhs.d, rhs.n_cov, rhs.nn_in, rhs.nn_out, len(rhs.scalar_names), rhs.series_param is not None,
                             tuple(sc0))


def make(ude, name, rhs, sc0, loss, solver, **kw):
    register(ude, name, rhs, sc0)
    prob, theta = cases.make_case(name, loss, solver, **kw)
    prob.user_rhs = rhs
    return prob, theta


def numpy_loss_grad(prob, theta):
    """loss and complex-step gradient of model 0 from the NumPy twin (oracle/ude_oracle_np.py)"""
    from oracle import ude_oracle_np as onp
    L = onp.loss(prob, np.asarray(theta, dtype=np.float64))
    g = onp.complex_step_grad(prob, theta)
    return L, np.array([g[k] for k in range(len(theta))])


def one_hot_model(ude, m=3, n_pts=8, hidden=6, distance=0.5, seed=3):
    """docs/src/MultipleTimeSeries.md:72-92: one network for m series, each told which series it is through a one-hot input"""
    import pandas as pd
    Y, ts = ude.synthetic.lotka_volterra_data(np.random.default_rng(0), m, n_pts, random_u0=True)
    df = pd.concat([pd.DataFrame({"time": ts + 0.1 * s, "series": f"s{s:02d}", "x1": Y[s, 0], "x2": Y[s, 1]}) for s in range(m)],
                   ignore_index=True)
    X = ude.one_hot_series_covariates(df, distance=distance)
    rhs = ude.expression_rhs(d=2, n_cov=m, nn_out=2, du=["y1", "y2"], name="ONE_HOT_NODE")
    _, nn = ude.SimpleNeuralNetwork(2 + m, 2, hidden=hidden, seed=seed)
    return ude.MultiCustomDerivatives(df, rhs, {"NN": nn}, X=X), nn, df
