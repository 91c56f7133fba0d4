"""Unscented-Kalman-filter marginal likelihood on the GPU (ude_ukf_loss_grad / ude_ukf_filter) against the NumPy restatement
(oracle/ukf_oracle_np.py): loss, complex-step gradient w.r.t. process model and process-noise factor, filtered states.

Tolerances.  With alpha = 1 the sigma-point weights are O(1) and the bar is the FP64 bar, 1e-10.  The reference's DEFAULT
alpha = 1e-3 gives weights Wm0 ~ -1e6, Wi ~ +5e5/L: the weighted sums cancel, and the error of ANY FP64 evaluation grows like
eps / alpha^2.  Measured against an extended-precision (np.longdouble) run of the twin (tools/ukf_alpha_conditioning.py,
profiles/ukf_alpha_r02.txt; tests/test_ukf_oracle.py::test_alpha_conditioning_extended_precision): at alpha = 1e-3 the FP64
twin is 6e-12 .. 4e-11 from the extended-precision loss and the CUDA path 1e-12 .. 3e-11 -- the kernel is as accurate as the
oracle it is checked against -- and the two FP64 gradients differ by up to 1.4e-9.  The bars below are 1e-9 (loss) and 5e-8
(gradient) for that case: 20-30x the measured spread, four orders tighter than round 1's 1e-5 / 1e-4.  The conditioning is
part of the reference's algorithm (src/UnscentedKalmanFilter.jl:2-8), not of this implementation."""
import numpy as np
import pytest

import cases
from oracle import ukf_oracle_np as ukf

pytestmark = pytest.mark.gpu


def _setup(rhs, hidden, lengths, solver, seed=0):
    prob, theta = cases.make_case(rhs, "cond_lik", solver, hidden=hidden, lengths=lengths, seed=seed)
    d = prob.d
    rng = np.random.default_rng(seed + 1)
    F = 0.3 * np.eye(d) + 0.05 * rng.normal(size=(d, d))
    Peta = 0.1 * np.eye(d) + 0.02 * np.ones((d, d))
    return prob, theta, F, Peta


@pytest.mark.parametrize("rhs,hidden,lengths", [("node_d2", 5, (7,)), ("node_d2", 5, (6, 3, 1, 4)), ("lv_ude_abs", 4, (6,)),
                                                ("node_d2", 5, (2,)), ("node_d4x1", 6, (2, 2, 1)),      # ONE filter step
                                                ("node_d4x1", 6, (4, 3)), ("node_time_d3", 5, (5, 2)), ("nn_decay_d1", 3, (9,))])
@pytest.mark.parametrize("solver", [0, 1])
def test_ukf_parity_alpha_one(ude, rhs, hidden, lengths, solver):
    prob, theta, F, Peta = _setup(rhs, hidden, lengths, solver)
    ref = ukf.loss(prob, theta, F, Peta, alpha=1.0, reg_weight=1e-3, reg_mult=2.0)
    g_pm, g_F = ukf.complex_step_grad(prob, theta, F, Peta, alpha=1.0, reg_weight=1e-3, reg_mult=2.0)
    with ude.Engine(prob) as eng:
        L, g, gF = eng.ukf_loss_grad(theta, F, Peta, alpha=1.0, reg=2e-3)
        L2, _, _ = eng.ukf_loss_grad(theta, F, Peta, alpha=1.0, reg=2e-3, want_grad=False)
    n_u = prob.d * prob.n_cols
    assert abs(L - ref) < 1e-10 * abs(ref) and L2 == L
    assert np.all(g[:n_u] == 0.0)
    assert cases.relerr(g[n_u:], g_pm) < 1e-10
    assert cases.relerr(gF.reshape(-1, order="F"), g_F) < 1e-10


@pytest.mark.parametrize("rhs,hidden,lengths", [("node_time_d3", 5, (6, 3)), ("node_d4x1", 6, (5, 4)), ("node_d2", 5, (6, 2))])
def test_ukf_prediction_window(ude, rhs, hidden, lengths):
    """Default = the reference's window [times[t], times[t] + dt] (src/UnscentedKalmanFilter.jl:83, :93); the causal window
    [times[t-1], times[t]] is an option.  Both against the twin; they differ unless the right-hand side is autonomous.  The
    option change and a second ude_set_data on the same handle must rebuild the cached filter plan."""
    prob, theta, F, Peta = _setup(rhs, hidden, lengths, 0)
    n_u = prob.d * prob.n_cols
    out = {}
    with ude.Engine(prob) as eng:
        for causal in (False, True, False):
            ref = ukf.loss(prob, theta, F, Peta, alpha=1.0, causal=causal)
            g_pm, g_F = ukf.complex_step_grad(prob, theta, F, Peta, alpha=1.0, causal=causal)
            L, g, gF = eng.ukf_loss_grad(theta, F, Peta, alpha=1.0, causal=causal)
            assert abs(L - ref) < 1e-10 * abs(ref), (rhs, causal)
            assert cases.relerr(g[n_u:], g_pm) < 1e-10 and cases.relerr(gF.reshape(-1, order="F"), g_F) < 1e-10
            out[causal] = L
        assert (abs(out[True] - out[False]) < 1e-12 * abs(out[True])) == (rhs == "node_d2")
        # new data on the same handle (more series, other lengths): the plan of the first problem must not be replayed
        prob2, theta2, F2, Peta2 = _setup(rhs, hidden, tuple(lengths) + (7, 2), 0, seed=3)
        eng._check(eng.lib.ude_set_data(eng.handle, prob2.n_series, prob2.starts.ctypes.data, prob2.lengths.ctypes.data, None,
                                        prob2.n_cols, prob2.times.ctypes.data, prob2.data.ctypes.data))
        if prob2.n_cov:
            eng._check(eng.lib.ude_set_covariates(eng.handle, prob2.knot_off.ctypes.data, prob2.knot_t.ctypes.data, prob2.knot_x.ctypes.data))
        eng.problem = prob2
        ref2 = ukf.loss(prob2, theta2, F2, Peta2, alpha=1.0)
        L2, _, _ = eng.ukf_loss_grad(theta2, F2, Peta2, alpha=1.0)
        assert abs(L2 - ref2) < 1e-10 * abs(ref2)


def test_ukf_reference_default_alpha(ude):
    prob, theta, F, Peta = _setup("node_d2", 5, (8, 5), 0)
    ref = ukf.loss(prob, theta, F, Peta, alpha=1e-3)
    g_pm, g_F = ukf.complex_step_grad(prob, theta, F, Peta, alpha=1e-3)
    with ude.Engine(prob) as eng:
        L, g, gF = eng.ukf_loss_grad(theta, F, Peta)
    n_u = prob.d * prob.n_cols
    assert abs(L - ref) < 1e-9 * abs(ref)
    assert cases.relerr(g[n_u:], g_pm) < 5e-8
    assert cases.relerr(gF.reshape(-1, order="F"), g_F) < 5e-8
    ref_ld = float(ukf.loss(prob, theta.astype(np.longdouble), F, Peta, alpha=1e-3))      # extended precision: the CUDA value is
    assert abs(L - ref_ld) < 1e-9 * abs(ref_ld)                                           # as close to it as the FP64 twin is


def test_ukf_filter_states(ude):
    prob, theta, F, Peta = _setup("node_d2", 5, (7, 4, 1), 0)
    n_u = prob.d * prob.n_cols
    pm = theta[n_u:n_u + prob.n_pm]
    with ude.Engine(prob) as eng:
        xs, Ps = eng.ukf_filter(theta, F, Peta, alpha=1.0)
    for s in range(prob.n_series):
        _, xr, Pr = ukf.filter_series(prob, pm, s, F @ F.T, Peta, 1.0, 2.0, 0.0)
        c0, n = int(prob.starts[s]), int(prob.lengths[s])
        assert np.max(np.abs(xs[:, c0:c0 + n] - xr)) < 1e-11
        assert np.max(np.abs(Ps[:, :, c0:c0 + n] - Pr)) < 1e-11


def test_ukf_linear_dynamics_known_answer(ude):
    """zero network + linear decay: the unscented transform is exact, the loss is the textbook Kalman filter's."""
    from oracle import ude_oracle_np as onp
    prob, theta = cases.make_case("nn_decay_d1", "cond_lik", 0, hidden=3, lengths=(9,))
    n_u = prob.d * prob.n_cols
    theta[n_u + prob.off_w2: n_u + prob.off_w2 + prob.nn_out * prob.nn_hidden] = 0.0
    theta[n_u + prob.off_b2: n_u + prob.off_b2 + prob.nn_out] = 0.0
    pm = theta[n_u:]
    y, tt = prob.data[0], prob.times
    x, P, ll = y[0], 0.05, 0.0
    for t in range(1, len(y)):
        a = onp.flow(prob, pm, 0, 0, np.array([1.0]), tt[t - 1], tt[t])[0]
        m, P1 = a * x, a * a * P + 0.04
        S = P1 + 0.05
        ll += -0.5 * (y[t] - m) ** 2 / S - 0.5 * np.log(S) - 0.5 * np.log(2 * 3.14159)
        x, P = m + P1 / S * (y[t] - m), P1 - P1 / S * P1
    with ude.Engine(prob) as eng:
        L, _, _ = eng.ukf_loss_grad(theta, np.array([[0.2]]), np.array([[0.05]]), alpha=1.0)
    assert abs(L + ll) < 1e-12 * abs(ll)


def test_train_marginal_likelihood(ude):
    """train!(model, loss_function = "marginal likelihood") (test/UDETests.jl:45 only checks it runs): the loss must go down,
    uhat must hold the filtered states and the returned P_nu must be the fitted factor's square."""
    df = ude.LotkaVolterra(plot=False, datasize=20)
    m = ude.NODE(df, hidden_units=6, seed=2)
    out = ude.train_(m, "marginal likelihood", verbose=False, regularization_weight=1e-4,
                     loss_options=dict(observation_error=0.05, alpha=1.0), optim_options=dict(maxiter=25, step_size=0.02))
    Pnu, Px = out
    tr = m.extra["loss_trace"]
    assert tr[-1] < tr[0] and np.all(np.isfinite(tr))
    assert Pnu.shape == (2, 2) and np.all(np.linalg.eigvalsh(Pnu) > 0) and Px.shape == (2, 2, 20)
    assert np.allclose(m.uhat[:, 0], m.data[:, 0]) and np.all(np.isfinite(m.uhat))
    # multi-series model: regulariser counted n_series + 1 times
    Y, ts = ude.synthetic.lotka_volterra_data(np.random.default_rng(0), 3, 9, random_u0=True)
    import pandas as pd
    dfm = pd.concat([pd.DataFrame({"time": ts, "series": s, "x1": Y[s, 0], "x2": Y[s, 1]}) for s in range(3)], ignore_index=True)
    mm = ude.MultiNODE(dfm, hidden_units=5, seed=1)
    ude.train_(mm, "marginal likelihood", verbose=False, loss_options=dict(alpha=1.0), optim_options=dict(maxiter=5))
    assert mm.extra["loss_trace"][-1] < mm.extra["loss_trace"][0]


def test_ukf_graph_replay_equals_plain_launches(ude, monkeypatch):
    prob, theta, F, Peta = _setup("node_d2", 5, (9, 4, 6), 0)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("UDE_UKF_GRAPH", flag)
        with ude.Engine(prob) as eng:
            a = eng.ukf_loss_grad(theta, F, Peta, alpha=1.0)
            b = eng.ukf_loss_grad(theta, 1.1 * F, Peta, alpha=1.0)       # a second call replays the captured graph with new constants
            res.append((a, b))
    for i in range(2):
        assert res[0][i][0] == res[1][i][0]
        assert np.array_equal(res[0][i][1], res[1][i][1]) and np.array_equal(res[0][i][2], res[1][i][2])
    assert res[0][0][0] != res[0][1][0]


def test_kalman_filter_multi(ude):
    """kalman_filter!(MultiUDE, sigma2) (src/UnscentedKalmanFilter.jl:181-235): diagonal process noise p.^2, first loss equals
    the oracle's with F = diag(p) and the regulariser counted once per series."""
    import pandas as pd
    Y, ts = ude.synthetic.lotka_volterra_data(np.random.default_rng(0), 3, 8, random_u0=True)
    dfm = pd.concat([pd.DataFrame({"time": ts, "series": s, "x1": Y[s, 0], "x2": Y[s, 1]}) for s in range(3)], ignore_index=True)
    mm = ude.MultiNODE(dfm, hidden_units=4, seed=1, reg_weight=1e-3)
    theta0 = mm.parameters.copy()
    prob = ude.default_loss(mm)
    ref = ukf.loss(prob, theta0, np.diag([0.1, 0.1]), 0.1 * np.eye(2), alpha=1.0, reg_weight=float(prob.reg_weight[0]), reg_mult=3.0)
    Pn, Px = ude.kalman_filter_(mm, 0.1, alpha=1.0, maxiter=6, step_size=0.02)
    tr = mm.extra["loss_trace"]
    assert abs(tr[0] - ref) < 1e-10 * abs(ref)
    assert tr[-1] < tr[0] and Pn.shape == (2, 2) and Px.shape == (2, 2, 24) and np.all(np.isfinite(mm.uhat))


def test_ukf_against_committed_fixtures(ude):
    """the CUDA path against tests/golden/golden_ukf_twin.npz directly (no oracle evaluation at test time)."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden_ukf as mg
    gold = np.load(os.path.join(here, "golden", "golden_ukf_twin.npz"))
    for rhs, hidden, lengths, solver, alpha in mg.CASES:
        prob, theta, F, Peta = mg.setup(rhs, hidden, lengths, solver)
        key = f"{rhs}|{hidden}|{'-'.join(map(str, lengths))}|{solver}|{alpha}"
        with ude.Engine(prob) as eng:
            L, g, gF = eng.ukf_loss_grad(theta, F, Peta, alpha=alpha, reg=2e-3)
        tl, tg = (1e-10, 1e-10) if alpha == 1.0 else (1e-9, 5e-8)
        assert abs(L - gold[key + "|loss"][0]) < tl * abs(L)
        assert cases.relerr(g[prob.d * prob.n_cols:], gold[key + "|g_pm"]) < tg
        assert cases.relerr(gF.reshape(-1, order="F"), gold[key + "|g_F"]) < tg


@pytest.mark.parametrize("rhs,hidden,lengths,n_models", [("node_d2", 5, (7, 4, 3, 2, 6, 5), 3), ("node_time_d3", 4, (5, 3, 4), 3),
                                                         ("node_d4x1", 6, (4, 6, 5, 2), 2), ("lv_ude_abs", 4, (3, 1, 6, 6), 2)])
@pytest.mark.parametrize("solver", [0, 1])
def test_ukf_many_models(ude, rhs, hidden, lengths, n_models, solver):
    """Many-model handle (folds x ensemble members): every model is filtered in the same launches, with its own process model
    and its own process-noise factor; a model whose series are shorter drops out of the later filter steps.  Each model
    against the single-model twin (loss, complex-step gradients), and against the same model in a handle of its own."""
    prob, theta = cases.make_case(rhs, "cond_lik", solver, hidden=hidden, lengths=lengths, n_models=n_models, seed=3)
    d = prob.d
    rng = np.random.default_rng(11)
    Fs = np.stack([0.3 * np.eye(d) + 0.05 * rng.normal(size=(d, d)) for _ in range(n_models)])
    Peta = 0.1 * np.eye(d) + 0.02 * np.ones((d, d))
    with ude.Engine(prob) as eng:
        L, g, gF = eng.ukf_loss_grad(theta, Fs, Peta, alpha=1.0, reg=2e-3)
        L0, _, _ = eng.ukf_loss_grad(theta, Fs, Peta, alpha=1.0, reg=2e-3, want_grad=False)
        xs, Ps = eng.ukf_filter(theta, Fs, Peta, alpha=1.0)
    assert L.shape == (n_models,) and gF.shape == (n_models, d, d) and np.array_equal(L0, L)
    touched = np.zeros(prob.n_theta, dtype=bool)
    for m in range(n_models):
        ref = ukf.loss(prob, theta, Fs[m], Peta, alpha=1.0, reg_weight=1e-3, reg_mult=2.0, model=m)
        g_pm, g_F = ukf.complex_step_grad(prob, theta, Fs[m], Peta, alpha=1.0, reg_weight=1e-3, reg_mult=2.0, model=m)
        off = ukf._model_range(prob, m)[0]
        assert abs(L[m] - ref) < 1e-10 * abs(ref), m
        assert cases.relerr(g[off:off + prob.n_pm], g_pm) < 1e-10, m
        assert cases.relerr(gF[m].reshape(-1, order="F"), g_F) < 1e-10, m
        touched[off:off + prob.n_pm] = True
        for s in range(int(prob.model_series0[m]), int(prob.model_series0[m + 1])):
            pm = theta[off:off + prob.n_pm]
            _, xr, Pr = ukf.filter_series(prob, pm, s, Fs[m] @ Fs[m].T, Peta, 1.0, 2.0, 0.0)
            c0, n = int(prob.starts[s]), int(prob.lengths[s])
            assert cases.relerr(xs[:, c0:c0 + n], xr) < 1e-10
            if n > 1:
                assert cases.relerr(Ps[:, :, c0:c0 + n], Pr) < 1e-10
    assert np.all(g[~touched] == 0.0)


def test_ukf_many_models_shared_factor_equals_single_handles(ude):
    """one (d, d) factor broadcast to every model; each model equals the same data in a handle of its own to round-off"""
    prob, theta = cases.make_case("node_d2", "cond_lik", 0, hidden=5, lengths=(6, 4, 5, 7), n_models=2, seed=5)
    d = prob.d
    F = 0.25 * np.eye(d); Peta = 0.1 * np.eye(d)
    with ude.Engine(prob) as eng:
        L, g, gF = eng.ukf_loss_grad(theta, F, Peta, alpha=1.0)
    for m in range(2):
        sub, th = cases.make_case("node_d2", "cond_lik", 0, hidden=5, lengths=(6, 4, 5, 7)[2 * m:2 * m + 2], seed=5)
        s0, s1 = int(prob.model_series0[m]), int(prob.model_series0[m + 1])
        c0, c1 = int(prob.model_col0[m]), int(prob.model_col0[m + 1])
        sub.data = np.asfortranarray(np.asarray(prob.data)[:, c0:c1]); sub.times = np.asarray(prob.times)[c0:c1].copy()
        sub.finalize()
        lo = int(prob.theta_base[m]); hi = lo + d * (c1 - c0) + prob.n_pm
        with ude.Engine(sub) as eng:
            L1, g1, gF1 = eng.ukf_loss_grad(theta[lo:hi], F, Peta, alpha=1.0)
        assert abs(L[m] - L1) <= 1e-12 * abs(L1)
        assert cases.relerr(g[lo:hi], g1) < 1e-11 and cases.relerr(gF[m], gF1) < 1e-11


def test_marginal_likelihood_folds_batched_equal_sequential(ude):
    """leave-future-out folds trained on the UKF likelihood in ONE many-model handle (marginal_likelihood_many_, the
    "marginal likelihood" branch of leave_future_out_batched) equal the folds trained one handle at a time."""
    df = ude.LotkaVolterra(plot=False, datasize=16)
    base = ude.NODE(df, hidden_units=5, seed=3)
    lo = dict(observation_error=0.05, alpha=1.0); oo = dict(maxiter=6, step_size=0.02)
    models, train, test, forecasts, trace = ude.leave_future_out_batched(base, 3, loss_function="marginal likelihood",
                                                                         regularization_weight=1e-4, loss_options=lo, optim_options=oo)
    assert len(models) == 3 and trace.shape == (6, 3) and np.all(np.isfinite(trace)) and len(forecasts) == 3
    for mi, tr in enumerate(train):
        one = base.constructor(tr)
        Pnu, Px = ude.train_(one, "marginal likelihood", verbose=False, regularization_weight=1e-4, loss_options=lo, optim_options=oo)
        assert cases.relerr(one.extra["loss_trace"], trace[:, mi]) < 1e-10
        assert cases.relerr(models[mi].parameters, one.parameters) < 1e-8
