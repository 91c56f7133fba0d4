"""The CPU oracle against (i) the committed golden vectors of the independent NumPy twin, (ii) complex-step derivatives,
(iii) closed forms, (iv) SciPy's adaptive solver (the reference integrates with adaptive Tsit5 at abstol=reltol=1e-6:
"trajectories within solver tolerance").  PARITY UNPINNED: no reference-produced vectors exist (see oracle/ude_oracle.h)."""
import os

import numpy as np
import pytest

import cases
from oracle import oracle, ude_oracle_np as onp

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_np_twin.npz"))
KEYS = sorted({k.rsplit("|", 1)[0] for k in GOLD.files})


@pytest.mark.parametrize("key", KEYS)
def test_oracle_matches_golden(key):
    rhs, loss, solver = key.split("|")
    prob, theta = cases.make_case(rhs, loss, int(solver), lengths=(6, 1, 9, 2), skip=(loss in ("cond_lik", "mse")))
    L, g = oracle.loss_grad(prob, theta)
    assert abs(L[0] - GOLD[key + "|loss"][0]) <= 1e-12 * abs(L[0])
    idx = GOLD[key + "|idx"]
    assert np.max(np.abs(g[idx] - GOLD[key + "|grad"])) <= 1e-11 * np.max(np.abs(g))


@pytest.mark.parametrize("loss", cases.LOSSES)
def test_oracle_gradient_complex_step(loss):
    prob, theta = cases.make_case("node_d4x1", loss, 1, lengths=(5, 3, 7), hidden=6)
    L, g = oracle.loss_grad(prob, theta)
    idx = np.arange(0, prob.n_theta, 3)
    cs = onp.complex_step_grad(prob, theta, 0, idx)
    assert max(abs(cs[k] - g[k]) for k in idx) <= 1e-12 * np.max(np.abs(g))


def test_oracle_threads_and_models():
    prob, theta = cases.make_case("node_time_d3", "cond_lik", 0, lengths=(9, 8, 7, 6, 5, 4), n_models=3)
    L1, g1 = oracle.loss_grad(prob, theta, threads=1)
    L4, g4 = oracle.loss_grad(prob, theta, threads=4)
    assert np.allclose(L1, L4, rtol=1e-13) and cases.relerr(g4, g1) < 1e-13
    assert np.allclose(L1, np.real(onp.loss(prob, theta)), rtol=1e-12)
    # loss-only call agrees and steps are counted like the host
    L0, g0 = oracle.loss_grad(prob, theta, want_grad=False)
    assert g0 is None and np.allclose(L0, L1, rtol=1e-14)
    assert oracle.count_steps(prob) == prob.count_steps()


def test_zero_network_is_exponential_decay():
    """du = NN(u) - m u with all NN weights zero is u' = -m u: u(t) = u0 exp(-m t) (cf. test/CrossValidation.jl:9-11)."""
    prob, theta = cases.make_case("nn_decay_d1", "shoot_theta", 0, lengths=(12,), substeps=32, reg=0.0)
    pm = prob.pm_view(theta, 0); pm[:] = 0.0; pm[prob.off_scalar[0]] = 0.7
    traj = oracle.forward(prob, theta)
    u0 = prob.uhat_view(theta, 0)[0, 0]
    exact = u0 * np.exp(-0.7 * (prob.times - prob.times[0]))
    assert np.max(np.abs(traj[0] - exact)) < 1e-10
    # Tsit5 (5th order) with fewer steps
    prob.solver = 1; prob.substeps = 8
    assert np.max(np.abs(oracle.forward(prob, theta)[0] - exact)) < 1e-10


def test_fixed_step_within_adaptive_solver_tolerance():
    """"Trajectories within solver tolerance": the reference integrates with adaptive Tsit5 at abstol = reltol = 1e-6
    (src/ProcessModels.jl:82-83).  Against a tight DOP853 solve (1e-12) of the same C3-shaped model, fixed-step RK4/Tsit5
    with 4 substeps per observation interval must be (a) within 1e-7 and (b) closer than an adaptive solve at 1e-6 is."""
    from scipy.integrate import solve_ivp
    import ude_b200 as ude
    prob, theta = ude.synthetic.config_c3(n_series=3)
    pm = prob.pm_view(theta, 0)
    prob.loss_kind = ude.problem.LOSS_SHOOT; prob.shoot_from_theta = False
    for solver in (0, 1):
        prob.solver = solver
        traj = oracle.forward(prob, theta)
        for s in range(3):
            c0, n = int(prob.starts[s]), int(prob.lengths[s])
            f = lambda t, u: np.real(onp.rhs(prob, pm, 0, s, u, t))
            span, te = (prob.times[c0], prob.times[c0 + n - 1]), prob.times[c0:c0 + n]
            tight = solve_ivp(f, span, prob.data[:, c0], t_eval=te, rtol=1e-12, atol=1e-12, method="DOP853", max_step=0.02)
            loose = solve_ivp(f, span, prob.data[:, c0], t_eval=te, rtol=1e-6, atol=1e-6, method="RK45")
            err_fixed = np.max(np.abs(traj[:, c0:c0 + n] - tight.y))
            err_adaptive = np.max(np.abs(loose.y - tight.y))
            assert err_fixed < 1e-7 and err_fixed < err_adaptive, (err_fixed, err_adaptive)


def test_multiple_shooting_block_structure():
    """Blocks start at t = 1, 1+pl, ... ; a series of pl*k + 1 points ends with a single-point block (App. A.4/A.5)."""
    import ude_b200 as ude
    prob, theta = cases.make_case("node_d2", "multi_shoot", 0, lengths=(11, 10, 1, 6), pred_length=5)
    assert prob.count_segments() == 3 + 2 + 1 + 2
    assert prob.count_steps() == (10 + 9 + 0 + 5) * prob.substeps
    out = oracle.forward(prob, theta)
    uh = prob.uhat_view(theta, 0)
    for c in (0, 5, 10, 11, 16, 21, 22, 27):      # block starts keep their uhat value (no pre-roll here)
        assert np.array_equal(out[:, c], uh[:, c])
