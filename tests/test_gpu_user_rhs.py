"""User-written right-hand sides (expressions -> generated trait -> nvcc plugin) on the GPU, through the C ABI.

Bars: 1e-10 against the NumPy twin (loss; complex-step gradient, element-wise per gradient leaf as in test_gpu_parity.py), and
against the hand-written catalog kernel where the expressions restate a catalog entry."""
import numpy as np
import pytest

import cases
import user_rhs_cases as urc
from oracle import oracle, ukf_oracle_np as ukf

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _check_np(prob, theta, ude):
    L_ref, g_ref = urc.numpy_loss_grad(prob, theta)
    with ude.Engine(prob) as eng:
        assert f"user{prob.rhs_kind}" in eng.kernel_name()
        L, g = eng.loss_grad(theta)
        L2, _ = eng.loss_grad(theta, want_grad=False)
    assert np.all(np.isfinite(g))
    assert abs(L[0] - L_ref[0]) < TOL * abs(L_ref[0]) and abs(L2[0] - L_ref[0]) < TOL * abs(L_ref[0])
    worst = cases.grad_mismatch(prob, g, g_ref, TOL)
    assert worst[0] <= 1.0, worst


@pytest.mark.parametrize("loss", ["cond_lik", "mse", "multi_shoot", "multi_shoot_preroll", "shoot_theta", "shoot_data", "deriv_match"])
@pytest.mark.parametrize("solver", [0, 1])
def test_expressions_restating_a_catalog_entry_equal_the_catalog_kernel(ude, loss, solver):
    rhs = urc.lv_twin(ude)
    kw = dict(hidden=10, lengths=(9, 1, 14, 2, 6), skip=(loss in ("cond_lik", "mse")))
    prob_u, theta = urc.make(ude, "user_lv", rhs, (1.0, 0.5, 0.5), loss, solver, **kw)
    prob_c, theta_c = cases.make_case("lv_ude", loss, solver, **kw)
    assert np.array_equal(theta, theta_c)
    L_ref, g_ref = oracle.loss_grad(prob_c, theta)
    with ude.Engine(prob_u) as eng:
        assert f"user{rhs.kind}" in eng.kernel_name()
        Lu, gu = eng.loss_grad(theta)
    with ude.Engine(prob_c) as eng:
        Lc, gc = eng.loss_grad(theta)
    assert abs(Lu[0] - Lc[0]) <= 1e-13 * abs(Lc[0]) and cases.relerr(gu, gc) < 1e-12
    assert abs(Lu[0] - L_ref[0]) < TOL * abs(L_ref[0])
    worst = cases.grad_mismatch(prob_c, gu, g_ref, TOL)
    assert worst[0] <= 1.0, worst


@pytest.mark.parametrize("loss,solver,n_models", [("cond_lik", 0, 1), ("multi_shoot", 1, 1), ("shoot_theta", 0, 1), ("mse", 0, 2)])
def test_per_series_parameter(ude, loss, solver, n_models):
    """`series_param="r"`: one trainable parameter per series (MultiCustomDerivatives' p.r[i]); the expression form of the
    catalog's per-series growth model against the hand-written kernel and the C oracle, also on a many-model handle"""
    rhs = urc.growth_twin(ude)
    kw = dict(hidden=4, lengths=(5, 3, 4, 6), n_models=n_models)
    prob_u, theta = urc.make(ude, "user_growth", rhs, (), loss, solver, **kw)
    prob_c, theta_c = cases.make_case("series_growth_d1", loss, solver, **kw)
    assert np.array_equal(theta, theta_c)
    L_ref, g_ref = oracle.loss_grad(prob_c, theta)
    with ude.Engine(prob_u) as eng:
        assert f"user{rhs.kind}" in eng.kernel_name()
        Lu, gu = eng.loss_grad(theta)
    with ude.Engine(prob_c) as eng:
        Lc, gc = eng.loss_grad(theta)
    assert np.max(np.abs(Lu - Lc) / np.abs(Lc)) <= 1e-13 and cases.relerr(gu, gc) < 1e-12
    assert np.max(np.abs(Lu - L_ref) / np.abs(L_ref)) < TOL
    worst = cases.grad_mismatch(prob_c, gu, g_ref, TOL)
    assert worst[0] <= 1.0, worst


@pytest.mark.parametrize("loss,solver,hidden", [("cond_lik", 0, 6), ("multi_shoot", 1, 6), ("deriv_match", 0, 5), ("shoot_theta", 0, 12),
                                                ("mse", 1, 20)])
def test_right_hand_side_outside_the_catalog(ude, loss, solver, hidden):
    """three states, covariate, time input, two network outputs, abs / exp / rational terms: against the NumPy twin"""
    rhs = urc.epidemic(ude)
    prob, theta = urc.make(ude, "user_epi", rhs, (0.9, 0.4, 2.0), loss, solver, hidden=hidden, lengths=(5, 3, 1, 4), substeps=2)
    _check_np(prob, theta, ude)


def test_user_rhs_through_the_host_api_and_the_filter(ude):
    """CustomDerivatives(data, expression_rhs, parameters) trains; forward predictions and the UKF run on the generated kernels"""
    df = ude.LotkaVolterra(plot=False, datasize=15)
    rhs = urc.lv_twin(ude)
    _, nn = ude.SimpleNeuralNetwork(2, 1, hidden=6, seed=4)
    m = ude.CustomDerivatives(df, rhs, {"NN": nn, "r": 0.5, "m": 0.4, "theta": 0.6})
    ude.train_(m, "conditional likelihood", verbose=False, optim_options=dict(maxiter=30, step_size=0.02))
    tr = m.extra["loss_trace"] if "loss_trace" in m.extra else None
    assert tr is None or tr[-1] < tr[0]
    inits, obs, preds = ude.predictions(m, df.iloc[3:9])
    assert preds.shape == obs.shape and np.all(np.isfinite(preds))
    # same model through the catalog entry: identical parameters after the same number of iterations
    mc = ude.CustomDerivatives(df, ude.LV_UDE, {"NN": nn, "r": 0.5, "m": 0.4, "theta": 0.6})
    ude.train_(mc, "conditional likelihood", verbose=False, optim_options=dict(maxiter=30, step_size=0.02))
    assert cases.relerr(m.parameters, mc.parameters) < 1e-9
    # UKF marginal likelihood on the generated kernels against the NumPy twin
    prob, theta = urc.make(ude, "user_lv", rhs, (1.0, 0.5, 0.5), "cond_lik", 0, hidden=5, lengths=(6, 3))
    F, Peta = 0.3 * np.eye(2), 0.1 * np.eye(2)
    ref = ukf.loss(prob, theta, F, Peta, alpha=1.0)
    g_pm, g_F = ukf.complex_step_grad(prob, theta, F, Peta, alpha=1.0)
    with ude.Engine(prob) as eng:
        L, g, gF = eng.ukf_loss_grad(theta, F, Peta, alpha=1.0)
    n_u = prob.d * prob.n_cols
    assert abs(L - ref) < TOL * abs(ref) and cases.relerr(g[n_u:], g_pm) < TOL and cases.relerr(gF.reshape(-1, order="F"), g_F) < TOL


@pytest.mark.parametrize("m,hidden", [(3, 6), (10, 10)])
def test_one_hot_series_network_on_the_gpu(ude, m, hidden):
    """docs/src/MultipleTimeSeries.md:72-92 (m = 10 series, hidden 10 in the docs): one-hot series input through generated kernels
    with n_cov = m constant covariates; loss and gradient against the NumPy twin, then a short training run"""
    model, nn, df = urc.one_hot_model(ude, m=m, n_pts=6, hidden=hidden)
    prob = ude.build_problem(model, "conditional likelihood")
    theta = model.parameters + 0.01 * np.random.default_rng(1).normal(size=model.parameters.shape)
    L_ref, g_ref = urc.numpy_loss_grad(prob, theta)
    with ude.Engine(prob) as eng:
        L, g = eng.loss_grad(theta)
    assert abs(L[0] - L_ref[0]) < TOL * abs(L_ref[0])
    worst = cases.grad_mismatch(prob, g, g_ref, TOL)
    assert worst[0] <= 1.0, worst
    ude.train_(model, "conditional likelihood", verbose=False, optim_options=dict(maxiter=40, step_size=0.02))
    assert np.all(np.isfinite(model.parameters))


def test_user_rhs_fp32_mode(ude):
    """the optional FP32 mode on generated kernels (the same trait text compiled in the float copy of the kernel family): bar 1e-4"""
    rhs = urc.lv_twin(ude)
    for loss, solver in (("cond_lik", 0), ("multi_shoot", 0)):
        prob_u, theta = urc.make(ude, "user_lv", rhs, (1.0, 0.5, 0.5), loss, solver, hidden=10, lengths=(9, 1, 14, 2, 6))
        prob_c, _ = cases.make_case("lv_ude", loss, solver, hidden=10, lengths=(9, 1, 14, 2, 6))
        L_ref, g_ref = oracle.loss_grad(prob_c, theta)
        with ude.Engine(prob_u, dtype=1) as eng:
            assert eng.kernel_name().startswith("f32_user")
            L, g = eng.loss_grad(theta)
        assert abs(L[0] - L_ref[0]) < 1e-4 * abs(L_ref[0]) and cases.relerr(g, g_ref) < 1e-4
    ep = urc.epidemic(ude)
    prob, theta = urc.make(ude, "user_epi", ep, (0.9, 0.4, 2.0), "multi_shoot", 0, hidden=6, lengths=(5, 3, 1, 4), substeps=2)
    L_ref, g_ref = urc.numpy_loss_grad(prob, theta)
    with ude.Engine(prob, dtype=1) as eng:
        L, g = eng.loss_grad(theta)
    assert abs(L[0] - L_ref[0]) < 1e-4 * abs(L_ref[0]) and cases.relerr(g, g_ref) < 1e-4


def test_every_training_route_accepts_a_generated_right_hand_side(ude, tmp_path):
    """train!'s loss functions and optimisers, the CV drivers and parameter I/O with an expression right-hand side: each route
    must run on the generated kernels and agree with the same model on the hand-written catalog kernel"""
    import pandas as pd
    df = ude.LotkaVolterra(plot=False, datasize=18)
    rhs = urc.lv_twin(ude)
    _, nn = ude.SimpleNeuralNetwork(2, 1, hidden=6, seed=4)
    init = {"NN": nn, "r": 0.5, "m": 0.4, "theta": 0.6}
    routes = [("spline gradient matching", "ADAM", dict(maxiter=15)), ("neural gradient matching", "ADAM", dict(maxiter=10)),
              ("derivative matching", "BFGS", dict(maxiter=5)), ("shooting", "ADAM", dict(maxiter=8, step_size=0.01)),
              ("multiple shooting", "ADAM", dict(maxiter=8, step_size=0.01)), ("conditional likelihood", "BFGS", dict(maxiter=4)),
              ("marginal likelihood", "ADAM", dict(maxiter=5))]
    for loss, opt, oo in routes:
        a = ude.CustomDerivatives(df, rhs, init)
        b = ude.CustomDerivatives(df, ude.LV_UDE, init)
        lo = dict(alpha=1.0) if loss == "marginal likelihood" else None
        ude.train_(a, loss, optimizer=opt, verbose=False, loss_options=lo, optim_options=oo)
        ude.train_(b, loss, optimizer=opt, verbose=False, loss_options=lo, optim_options=oo)
        assert np.all(np.isfinite(a.parameters)), loss
        assert cases.relerr(a.parameters, b.parameters) < 1e-7, (loss, cases.relerr(a.parameters, b.parameters))
    f = tmp_path / "pm.json"
    ude.save_model_parameters(a, str(f))
    c = ude.CustomDerivatives(df, rhs, init)
    ude.load_model_parameters_(c, str(f))
    assert np.array_equal(c.process_model, a.process_model)
    tr, te, fc = ude.leave_future_out(a, lambda mi: ude.train_(mi, "conditional likelihood", verbose=False, optim_options=dict(maxiter=3)), 2)
    assert len(fc) == 2 and np.all(np.isfinite(fc[0][["x1", "x2"]].to_numpy()))
    # multi-series model with a per-series parameter: leave-site-out refits on the other sites
    rng = np.random.default_rng(2)
    ts = np.linspace(0.0, 3.0, 10)
    dfm = pd.concat([pd.DataFrame({"time": ts, "series": s, "x": 0.4 + 0.1 * s + 0.05 * np.sin(ts) + 0.01 * rng.normal(size=10)}) for s in (1, 2, 3)],
                    ignore_index=True)
    _, nn1 = ude.SimpleNeuralNetwork(1, 1, hidden=4, seed=1)
    mm = ude.MultiCustomDerivatives(dfm, urc.growth_twin(ude), {"NN": nn1, "r": np.array([0.1, 0.2, 0.3])})
    score = ude.leave_site_out(mm, lambda mi: ude.train_(mi, "conditional likelihood", verbose=False, optim_options=dict(maxiter=3)))
    assert np.isfinite(score)
