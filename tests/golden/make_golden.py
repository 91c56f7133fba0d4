"""Generates tests/golden/*.npz -- run from the repo root:  python tests/golden/make_golden.py

PARITY UNPINNED: the Julia reference cannot be executed here (no julia binary), and its own tests hold
no numeric expectations, so these fixtures are NOT reference outputs.  They are produced by the NumPy twin
(oracle/ude_oracle_np.py: an independently written forward evaluation of the reference's loss formulas) with
complex-step derivatives, and pin the C oracle and the CUDA path against silent drift between rounds.
Inputs are regenerated from the seed inside tests/cases.py; only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import ude_oracle_np as onp  # noqa: E402

GOLDEN = [(rhs, loss, solver) for rhs in cases.RHS_SPECS for loss in cases.LOSSES for solver in (0, 1)
          if (hash_ := (list(cases.RHS_SPECS).index(rhs) + cases.LOSSES.index(loss) + solver) % 3 == 0)]


def main():
    out = {}
    for rhs, loss, solver in GOLDEN:
        prob, theta = cases.make_case(rhs, loss, solver, lengths=(6, 1, 9, 2), skip=(loss in ("cond_lik", "mse")))
        L = onp.loss(prob, theta)
        rng = np.random.default_rng(11)
        idx = np.unique(np.concatenate([rng.choice(prob.n_theta, 8, replace=False),
                                        np.arange(prob.pm_base(0), prob.n_theta)[::5], [prob.n_theta - 1]]))
        cs = onp.complex_step_grad(prob, theta, 0, idx)
        key = f"{rhs}|{loss}|{solver}"
        out[key + "|loss"] = np.real(L)
        out[key + "|idx"] = idx
        out[key + "|grad"] = np.array([cs[k] for k in idx])
        print(key, float(np.real(L[0])))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_np_twin.npz"), **out)


if __name__ == "__main__":
    main()
