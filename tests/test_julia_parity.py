"""Pins the CPU oracle to the reference's own Julia implementation -- WHEN tests/golden/julia_parity.json exists.

That file is written by `baseline/julia_parity.jl` (run where Julia + UniversalDiffEq.jl are installed; this image has
neither, so the file is absent and these tests SKIP: parity stays "unpinned").  The inputs the Julia script
reads (tests/golden/julia_cases/) are committed and are regenerated bit-identically here from the host API, so the day the
JSON is produced the oracle -- and through tests/test_gpu_parity.py the CUDA path -- is compared with
OrdinaryDiffEq's fixed-step RK4 / Tsit5 steppers driven through the reference's model objects at the north-star bar."""
import json
import os
import sys

import numpy as np
import pytest

import cases
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_julia_inputs as mj  # noqa: E402

GOLD = os.path.join(HERE, "golden", "julia_parity.json")


def test_julia_inputs_are_reproducible():
    """the committed inputs of the Julia script are exactly what the host API builds today (so a JSON produced from them pins
    the current oracle, not a stale problem)"""
    for name, model_kind, loss, solver, hidden in mj.CASES:
        m, prob, theta, df, X = mj.build(name, model_kind, loss, solver, hidden)
        d = os.path.join(mj.OUT, name)
        th = np.loadtxt(os.path.join(d, "theta.txt"))
        assert th.shape[0] == prob.n_theta and np.array_equal(th, theta), name
        import pandas as pd
        assert np.allclose(pd.read_csv(os.path.join(d, "data.csv"), float_precision="round_trip").to_numpy(dtype=float), df.to_numpy(dtype=float), rtol=0, atol=0)
        meta = dict(l.strip().split(" = ") for l in open(os.path.join(d, "meta.txt")))
        assert int(meta["n_theta"]) == prob.n_theta and int(meta["nn_in"]) == prob.nn_in


@pytest.mark.skipif(not os.path.exists(GOLD), reason="tests/golden/julia_parity.json absent: baseline/julia_parity.jl has not been run (no Julia here)")
def test_oracle_matches_julia_reference():
    gold = json.load(open(GOLD))
    for name, model_kind, loss, solver, hidden in mj.CASES:
        m, prob, theta, df, X = mj.build(name, model_kind, loss, solver, hidden)
        prob.reg_weight = np.zeros_like(prob.reg_weight)          # the Julia script evaluates the data-misfit terms only
        L, g = oracle.loss_grad(prob, theta)
        ref = gold[name]
        assert abs(L[0] - ref["loss"]) < 1e-10 * abs(ref["loss"]), (name, L[0], ref["loss"])
        assert cases.grad_mismatch(prob, g, np.asarray(ref["grad"]), rtol=1e-10)[0] <= 1.0, name
        if ref.get("reference_adaptive_loss") is not None:        # adaptive Tsit5 at 1e-6: the reference's own closure
            assert abs(L[0] - ref["reference_adaptive_loss"]) < 1e-3 * abs(L[0]), name
