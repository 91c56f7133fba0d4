"""Writes tests/golden/julia_cases/<case>/{data.csv, covariates.csv, theta.txt, meta.txt}: the inputs of
baseline/julia_parity.jl, generated from the SAME host-API objects tests/test_julia_parity.py rebuilds (seeded NumPy).
Run from the repo root:  python tests/golden/make_julia_inputs.py"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ude_b200 as ude  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "julia_cases")
CASES = [  # name, model, loss, solver, hidden
    ("node_condlik_rk4", "NODE", "conditional likelihood", "rk4", 6),
    ("node_multishoot_tsit5", "NODE", "multiple shooting", "tsit5", 6),
    ("node_shooting_rk4", "NODE", "shooting", "rk4", 5),
    ("nodex_multishoot_rk4", "NODE_X", "multiple shooting", "rk4", 6),
    ("multinodex_multishoot_rk4", "MultiNODE_X", "multiple shooting", "rk4", 8),
    ("multinodex_condlik_tsit5", "MultiNODE_X", "conditional likelihood", "tsit5", 8),
]
SUBSTEPS, PRED_LENGTH = 4, 5


def frames(model_kind, rng):
    """(data frame, covariate frame or None) in the column layout the reference's constructors accept."""
    if model_kind == "MultiNODE_X":
        Y, ts = ude.synthetic.lotka_volterra_data(rng, 3, 12, T=2.0, random_u0=True)
        lens = (12, 9, 7)
        df = pd.concat([pd.DataFrame({"time": ts[:n], "series": s + 1, "x1": Y[s, 0, :n], "x2": Y[s, 1, :n]}) for s, n in enumerate(lens)],
                       ignore_index=True)
        X = pd.concat([pd.DataFrame({"time": np.linspace(-0.1, 2.2, 8), "series": s + 1, "X1": rng.normal(size=8)}) for s in range(3)],
                      ignore_index=True)
        return df, X
    Y, ts = ude.synthetic.lotka_volterra_data(rng, 1, 14, T=2.0)
    df = pd.DataFrame({"time": ts, "x1": Y[0, 0], "x2": Y[0, 1]})
    X = pd.DataFrame({"time": np.linspace(-0.1, 2.2, 9), "X1": rng.normal(size=9)}) if model_kind == "NODE_X" else None
    return df, X


def build(name, model_kind, loss, solver, hidden):
    """-> (host model, Problem, theta) exactly as tests/test_julia_parity.py needs them"""
    rng = np.random.default_rng(sum(map(ord, name)))
    df, X = frames(model_kind, rng)
    ode = ude.CudaTsit5(SUBSTEPS) if solver == "tsit5" else ude.CudaRK4(SUBSTEPS)
    if model_kind == "MultiNODE_X":
        m = ude.MultiNODE(df, X, hidden_units=hidden, seed=1, ode_solver=ode)
    else:
        m = ude.NODE(df, X, hidden_units=hidden, seed=1, ode_solver=ode)
    if loss == "conditional likelihood":
        prob = ude.training.conditional_likelihood(m, observation_error=0.025, process_error=0.05)
    elif loss == "multiple shooting":
        prob = ude.training.multiple_shooting_loss(m, pred_length=PRED_LENGTH)
    else:
        prob = ude.training.shooting_loss(m)
    theta = m.parameters.copy()
    theta += 0.05 * rng.normal(size=theta.shape)
    n_u = prob.d * prob.n_cols
    theta[:n_u] = np.asarray(prob.data).reshape(-1, order="F") + 0.03 * rng.normal(size=n_u)
    return m, prob, theta, df, X


def main():
    for name, model_kind, loss, solver, hidden in CASES:
        m, prob, theta, df, X = build(name, model_kind, loss, solver, hidden)
        d = os.path.join(OUT, name)
        os.makedirs(d, exist_ok=True)
        df.to_csv(os.path.join(d, "data.csv"), index=False, float_format="%.17g")
        if X is not None:
            X.to_csv(os.path.join(d, "covariates.csv"), index=False, float_format="%.17g")
        np.savetxt(os.path.join(d, "theta.txt"), theta, fmt="%.17g")
        with open(os.path.join(d, "meta.txt"), "w") as f:
            f.write(f"model = {model_kind}\nloss = {loss}\nsolver = {solver}\nhidden = {hidden}\nnn_in = {prob.nn_in}\n"
                    f"substeps = {SUBSTEPS}\npred_length = {PRED_LENGTH}\nobservation_error = 0.025\nprocess_error = 0.05\n"
                    f"shoot_weight = {float(prob.shoot_weight[0]):.17g}\nn_theta = {prob.n_theta}\n")
        print(name, prob.n_theta)


if __name__ == "__main__":
    main()
