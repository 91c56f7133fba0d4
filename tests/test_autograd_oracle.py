"""A third, independent gradient for the CPU oracle: torch.float64 reverse-mode autograd through a
re-assembly of the reference's losses (oracle/ude_autograd_torch.py) against the hand-derived discrete adjoint of
oracle/ude_oracle.c -- every loss kind x every right-hand side of tests/cases.py, element-wise per gradient leaf at the
north-star bar (1e-10).  Together with the complex-step twin (tests/test_oracle.py) the oracle's gradient is now checked by
two methods that share no code with it or with each other.  PARITY stays UNPINNED against Julia."""
import numpy as np
import pytest

import cases
from oracle import oracle
from oracle import ude_autograd_torch as ag


@pytest.mark.parametrize("loss", cases.LOSSES)
@pytest.mark.parametrize("rhs", list(cases.RHS_SPECS))
def test_oracle_gradient_equals_torch_autograd(rhs, loss):
    hidden = 6 if rhs != "node_d8" else 9
    for solver in ((0, 1) if rhs in ("node_d4x1", "node_time_d3", "lv_ude_abs", "series_growth_d1", "node_d2") else (0,)):
        prob, theta = cases.make_case(rhs, loss, solver, hidden=hidden, lengths=(7, 1, 6, 2), substeps=2,
                                      skip=(loss in ("cond_lik", "mse")))
        L_ref, g_ref = oracle.loss_grad(prob, theta)
        L, g = ag.loss_grad(prob, theta)
        assert np.max(np.abs(L - L_ref) / np.abs(L_ref)) < 1e-12, (rhs, loss, solver, L, L_ref)
        worst = cases.grad_mismatch(prob, g, g_ref, rtol=1e-10)
        assert worst[0] <= 1.0, (rhs, loss, solver, worst)
        assert np.all(g[g_ref == 0.0] == 0.0)


def test_many_models_and_euler():
    """CV folds x members in one theta (independent blocks) and the explicit-Euler "solver" of spline gradient matching."""
    for loss in ("cond_lik", "mse", "multi_shoot", "shoot_theta", "deriv_match"):
        prob, theta = cases.make_case("node_time_d3", loss, 0, hidden=5, lengths=(5, 4, 3, 6, 2, 3), n_models=3, substeps=2)
        L_ref, g_ref = oracle.loss_grad(prob, theta)
        L, g = ag.loss_grad(prob, theta)
        assert np.max(np.abs(L - L_ref) / np.abs(L_ref)) < 1e-12
        assert cases.grad_mismatch(prob, g, g_ref, rtol=1e-10)[0] <= 1.0, loss
    prob, theta = cases.make_case("lv_ude", "mse", 2, hidden=5, lengths=(6, 3), substeps=1)
    L_ref, g_ref = oracle.loss_grad(prob, theta)
    L, g = ag.loss_grad(prob, theta)
    assert np.max(np.abs(L - L_ref) / np.abs(L_ref)) < 1e-12
    assert cases.grad_mismatch(prob, g, g_ref, rtol=1e-10)[0] <= 1.0
