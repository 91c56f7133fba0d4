"""Generates tests/golden/golden_ukf_twin.npz -- run from the repo root:  python tests/golden/make_golden_ukf.py

PARITY UNPINNED (see make_golden.py): these are outputs of oracle/ukf_oracle_np.py (NumPy restatement of
src/UnscentedKalmanFilter.jl with complex-step gradients), stored to pin the twin and the CUDA path against drift between
rounds.  Inputs are regenerated from the seeds inside tests/cases.py; only outputs are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import ukf_oracle_np as ukf  # noqa: E402

CASES = [("node_d2", 5, (7,), 0, 1.0), ("node_d2", 5, (6, 3, 1, 4), 1, 1.0), ("lv_ude_abs", 4, (6,), 0, 1.0),
         ("node_d4x1", 6, (4, 3), 0, 1.0), ("node_time_d3", 5, (5, 2), 1, 1.0), ("node_d2", 5, (8, 5), 0, 1e-3)]


def setup(rhs, hidden, lengths, solver, seed=0):
    prob, theta = cases.make_case(rhs, "cond_lik", solver, hidden=hidden, lengths=lengths, seed=seed)
    d = prob.d
    rng = np.random.default_rng(seed + 1)
    F = 0.3 * np.eye(d) + 0.05 * rng.normal(size=(d, d))
    Peta = 0.1 * np.eye(d) + 0.02 * np.ones((d, d))
    return prob, theta, F, Peta


def main():
    out = {}
    for rhs, hidden, lengths, solver, alpha in CASES:
        prob, theta, F, Peta = setup(rhs, hidden, lengths, solver)
        L = ukf.loss(prob, theta, F, Peta, alpha=alpha, reg_weight=1e-3, reg_mult=2.0)
        g_pm, g_F = ukf.complex_step_grad(prob, theta, F, Peta, alpha=alpha, reg_weight=1e-3, reg_mult=2.0)
        key = f"{rhs}|{hidden}|{'-'.join(map(str, lengths))}|{solver}|{alpha}"
        out[key + "|loss"] = np.array([L])
        out[key + "|g_pm"] = g_pm
        out[key + "|g_F"] = g_F
        print(key, L)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_ukf_twin.npz"), **out)


if __name__ == "__main__":
    main()
