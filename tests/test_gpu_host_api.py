"""train! / gradient_descent! / leave_future_out through the host-side mirror on the GPU (reference tests:
test/UDETests.jl:41-47, test/MultiUDE.jl:30-36, test/cross_validation.jl:36 -- smoke tests there; here every result is
compared with the same optimiser driven by oracle gradients)."""
import numpy as np
import pandas as pd
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def _lv_frame(ude, n=22, seed=3):
    Y, ts = ude.synthetic.lotka_volterra_data(np.random.default_rng(seed), 1, n)
    return pd.DataFrame({"time": ts, "x1": Y[0, 0], "x2": Y[0, 1]})


def _oracle_adam(prob, theta, step, iters):
    th = theta.copy(); m = np.zeros_like(th); v = np.zeros_like(th)
    for it in range(1, iters + 1):
        _, g = oracle.loss_grad(prob, th)
        m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
        th = th - step * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
    return th


@pytest.mark.parametrize("loss_function", ["conditional likelihood", "multiple shooting", "shooting", "derivative matching"])
@pytest.mark.parametrize("on_device", [True, False])
def test_train_matches_oracle_adam(ude, loss_function, on_device):
    df = _lv_frame(ude)
    _, p = ude.SimpleNeuralNetwork(2, 1, seed=123)
    m = ude.CustomDerivatives(df, ude.LV_UDE, {"NN": p, "r": 1.0, "m": 0.5, "theta": 0.5})        # test/UDETests.jl:20-31
    m.parameters[: 2 * m.N] = m.data.reshape(-1, order="F")
    theta0 = m.parameters.copy()
    prob = ude.build_problem(m, loss_function, 0.0)
    ude.train_(m, loss_function=loss_function, optimizer="ADAM", verbose=False, optim_options=dict(maxiter=4), on_device=on_device)
    ref = _oracle_adam(prob, theta0, ude.training.default_options(loss_function)["step_size"], 4)
    pmb = prob.pm_base(0)
    assert np.max(np.abs(m.parameters[pmb:] - ref[pmb:])) < 1e-8 * max(1.0, np.max(np.abs(ref[pmb:])))
    if loss_function == "conditional likelihood":
        assert np.max(np.abs(m.parameters[:pmb] - ref[:pmb])) < 1e-8
    elif loss_function in ("shooting", "multiple shooting"):      # uhat is overwritten with the fitted trajectories
        assert np.max(np.abs(m.uhat - oracle.forward(prob, ref))) < 1e-7


def test_multinode_default_loss_and_gradient_descent(ude):
    df = _lv_frame(ude, n=15)
    dfm = pd.concat([df.assign(series=1), df.iloc[:11].assign(series=2), df.iloc[:7].assign(series=3)])
    X = pd.DataFrame({"time": np.tile(np.linspace(0, 3, 9), 3), "series": np.repeat([1, 2, 3], 9), "X": np.random.default_rng(0).normal(size=27)})
    m = ude.MultiNODE(dfm, X, hidden_units=10, seed=1)
    theta0 = m.parameters.copy()
    prob = ude.default_loss(m)
    ude.gradient_descent_(m, step_size=0.05, maxiter=3)
    ref = _oracle_adam(prob, theta0, 0.05, 3)
    assert np.max(np.abs(m.parameters - ref)) < 1e-8


def test_leave_future_out_batched_equals_serial(ude):
    """k folds as ONE batched handle == k separate fits (src/CrossValidation.jl:17-45 runs them under Threads.@threads)."""
    df = _lv_frame(ude, n=20)
    _, p = ude.SimpleNeuralNetwork(2, 1, seed=123)
    m = ude.CustomDerivatives(df, ude.LV_UDE_ABS, {"NN": p, "r": 0.1, "k": 10.1, "m": 0.1, "theta": 0.1})
    opts = dict(maxiter=5, step_size=0.025)
    models, tr, te, fc, trace = ude.leave_future_out_batched(m, k=3, skip=1, loss_function="conditional likelihood", optim_options=opts)
    training = lambda mi: ude.train_(mi, loss_function="conditional likelihood", verbose=False, optim_options=opts)
    tr2, te2, fc2 = ude.leave_future_out(m, training, 3)
    assert [len(t) for t in tr] == [19, 18, 17] and [len(t) for t in te] == [1, 2, 3]
    for a, b in zip(fc, fc2):
        assert np.allclose(a.to_numpy(), b.to_numpy(), rtol=1e-9, atol=1e-12)
    by_h, _ = ude.summarize_leave_future_out(m, te, fc)
    assert len(by_h) == 3 and np.all(np.isfinite(by_h["mean_absolute_error"]))
    assert trace.shape == (5, 3) and np.all(np.diff(trace[:, 0]) < 0)        # the loss goes down


def test_forecast_and_predictions_match_oracle(ude):
    """forecast() (src/ModelTesting.jl:541-571 with the damping of src/ProcessModels.jl:2-25) and predictions()
    (:181-195) through the GPU against the same steps integrated by the oracle."""
    df = _lv_frame(ude, n=18)
    X = pd.DataFrame({"time": np.linspace(-0.5, 4.5, 11), "X": np.sin(np.linspace(0, 3, 11))})
    m = ude.NODE(df, X, hidden_units=10, seed=3)
    m.parameters[: 2 * m.N] = m.data.reshape(-1, order="F")
    prob = ude.default_loss(m)
    pred = ude.predictions(m)
    assert np.max(np.abs(pred - oracle.forward(prob, m.parameters))) < 1e-10
    t_new = m.times[-1] + np.array([0.1, 0.25, 0.6, 1.0])
    fc = ude.forecast(m, m.uhat[:, -1], m.times[-1], t_new)
    # the same recursion with oracle one-step predictions
    from oracle import ude_oracle_np as onp
    x = m.uhat[:, -1].copy(); t0 = m.times[-1]
    umax, umin, umean = m.uhat.max(axis=1), m.uhat.min(axis=1), m.uhat.mean(axis=1)
    for j, t1 in enumerate(t_new):
        p1 = np.real(onp.flow(prob, m.process_model, 0, 0, x, t0, t1))
        du = p1 - x
        w = np.exp(-0.5 / 0.25 ** 2 * ((x - umax) / (umax - umin)) ** 2); du = np.where(x > umax, w * du - 0.1 * (1 - w) * (x - umean), du)
        w = np.exp(-0.5 / 0.25 ** 2 * ((x - umin) / (umax - umin)) ** 2); du = np.where(x < umin, w * du - (1 - w) * 0.1 * (x - umean), du)
        x = x + du; t0 = t1
        assert np.max(np.abs(fc[j, 1:] - x)) < 1e-9 and fc[j, 0] == t1


def test_forecast_with_a_per_series_parameter(ude):
    """MultiCustomDerivatives with `p.r[i]` (docs/src/MultipleTimeSeries.md:27-61): forecast() and the leave-future-out forecasts
    chain one-interval mini-series; each of them must carry the growth rate of the series being forecast."""
    from oracle import ude_oracle_np as onp
    rng = np.random.default_rng(3)
    ts = np.linspace(0.0, 3.0, 12)
    dfm = pd.concat([pd.DataFrame({"time": ts, "series": s, "x": 0.3 + 0.2 * s + 0.1 * np.sin(ts + s) + 0.01 * rng.normal(size=12)}) for s in (1, 2, 3)],
                    ignore_index=True)
    _, nn = ude.SimpleNeuralNetwork(1, 1, hidden=5, seed=2)
    m = ude.MultiCustomDerivatives(dfm, ude.series_growth_rhs(1), {"NN": nn, "r": np.array([0.3, -0.2, 0.5])})
    m.parameters[: m.data.shape[1]] = m.data.reshape(-1, order="F")
    prob = ude.default_loss(m)
    for j in range(3):
        s0, ln = int(m.starts[j]), int(m.lengths[j])
        uh = m.uhat[:, s0:s0 + ln]
        t_new = m.times[s0 + ln - 1] + np.array([0.2, 0.5, 0.9])
        fc = ude.forecast(m, uh[:, -1], m.times[s0 + ln - 1], t_new, series=j)
        x = uh[:, -1].copy(); t0 = m.times[s0 + ln - 1]
        umax, umin, umean = uh.max(axis=1), uh.min(axis=1), uh.mean(axis=1)
        l, rho = m.extra.get("l", 0.25), m.extra.get("extrap_rho", 0.1)
        for k, t1 in enumerate(t_new):
            p1 = np.real(onp.flow(prob, m.process_model, j, j, x, t0, t1))
            du = p1 - x
            w = np.exp(-0.5 / l ** 2 * ((x - umax) / (umax - umin)) ** 2); du = np.where(x > umax, w * du - rho * (1 - w) * (x - umean), du)
            w = np.exp(-0.5 / l ** 2 * ((x - umin) / (umax - umin)) ** 2); du = np.where(x < umin, w * du - (1 - w) * rho * (x - umean), du)
            x = x + du; t0 = t1
            assert np.max(np.abs(fc[k, 1:] - x)) < 1e-9, (j, k)
    models, tr, te, fc, trace = ude.leave_future_out_batched(m, k=2, skip=1, loss_function="conditional likelihood", optim_options=dict(maxiter=3))
    assert len(fc) == 2 and np.all(np.isfinite(fc[0][["x"]].to_numpy()))


def test_multinode_leave_future_out_and_kfold(ude):
    df = _lv_frame(ude, n=14)
    dfm = pd.concat([df.assign(series="a"), df.iloc[:12].assign(series="b")])
    m = ude.MultiNODE(dfm, hidden_units=10)
    opts = dict(maxiter=3, step_size=0.05)
    models, tr, te, fc, trace = ude.leave_future_out_batched(m, k=2, skip=1, loss_function="multiple shooting", optim_options=opts)
    assert len(fc) == 2 and set(fc[0]["series"]) == {"a", "b"} and np.all(np.isfinite(fc[1][["x1", "x2"]].to_numpy()))
    m1 = ude.NODE(df, hidden_units=10)
    out = ude.interval_holdout_cv(m1, lambda mi, ts: ude.train_(mi, "conditional likelihood", verbose=False, t_skip=ts, optim_options=dict(maxiter=2)), k=3)
    assert len(out) == 3 and np.all(np.isfinite(out["mean_absolute_error"]))


def test_ensemble_times_folds_in_one_handle(ude):
    """BASELINE config C4 in miniature: members (different NN seeds) x leave-future-out folds = independent theta vectors
    trained together; every (member, fold) equals its own separate fit."""
    df = _lv_frame(ude, n=16)
    m = ude.NODE(df, hidden_units=10, seed=1)
    opts = dict(maxiter=4, step_size=0.05)
    models, tr, te, fc, trace = ude.leave_future_out_batched(m, k=2, skip=1, loss_function="conditional likelihood", optim_options=opts,
                                                              member_seeds=[1, 7, 9], do_forecast=False)
    assert len(models) == 6 and trace.shape == (4, 6)
    assert not np.allclose(models[0].process_model, models[2].process_model)          # different seeds
    for mi in (0, 3, 5):
        solo = ude.NODE(tr[mi], hidden_units=10, seed=[1, 7, 9][mi // 2])
        ude.train_(solo, "conditional likelihood", verbose=False, optim_options=opts)
        assert np.max(np.abs(solo.parameters - models[mi].parameters)) < 1e-9


def test_leave_future_out_predict_matches_oracle_steps(ude):
    """leave_future_out_predict (src/CrossValidation.jl:608-640; the reference's test only checks it runs,
    test/cross_validation.jl:62): every one-step change must equal the oracle's flow from the observed start."""
    from oracle import ude_oracle_np as onp
    df = ude.LotkaVolterra(plot=False, datasize=26)
    model = ude.NODE(df, hidden_units=10, seed=5)
    fitted = []

    def training(m):
        ude.train_(m, "conditional likelihood", verbose=False, optim_options=dict(maxiter=3, step_size=0.01))
        fitted.append(m)

    summary, raw = ude.leave_future_out_predict(model, training, 3, 2)
    assert list(summary.columns) == ["variable", "MSE", "MSE_SE", "MAE", "MAE_SE", "correlation", "count"]
    assert list(summary["count"]) == [6, 6] and len(raw) == 12 and len(fitted) == 2
    d = df.sort_values("time").reset_index(drop=True)
    n = len(d)
    for fold, m in enumerate(fitted):
        i = 3 * (fold + 1)
        test = d.iloc[n - i - 1: n - i + 3]
        prob = ude.default_loss(m)
        tt = test["time"].to_numpy(); Y = test[["x1", "x2"]].to_numpy().T
        for j in range(3):
            p1 = np.real(onp.flow(prob, m.process_model, 0, 0, Y[:, j], tt[j], tt[j + 1]))
            for v, name in enumerate(["x1", "x2"]):
                row = raw[(raw["time"] == tt[j + 1]) & (raw["variable"] == name)]
                assert len(row) == 1
                assert abs(row["predicted"].iloc[0] - (p1[v] - Y[v, j])) < 1e-10
                assert abs(row["observed"].iloc[0] - (Y[v, j + 1] - Y[v, j])) < 1e-14


def test_train_bfgs_reaches_lower_loss_than_start(ude):
    """train!(...; optimizer = "BFGS") (src/Optimizers.jl:506-521; docs/src/examples.md:180): same loss/gradient entry point."""
    m = ude.NODE(_lv_frame(ude, n=14), hidden_units=5, seed=4)
    ude.train_(m, "conditional likelihood", optimizer="BFGS", verbose=False, optim_options=dict(maxiter=15))
    tr = m.extra["loss_trace"].ravel()
    assert tr.min() < 0.5 * tr[0] and np.all(np.isfinite(m.parameters))
    with pytest.raises(ValueError):
        ude.train_(m, "conditional likelihood", optimizer="SGD")


def test_spline_gradient_matching_default_loss(ude):
    """train!(model) with its default loss, spline gradient matching (src/Optimizers.jl:354, LFC.jl:888-931; reference test:
    test/UDETests.jl:41 only checks it runs).  One evaluation against the formula written out with the NumPy RHS, the
    gradient against the oracle-driven twin, then a short training run."""
    from oracle import ude_oracle_np as onp
    df = ude.LotkaVolterra(plot=False, datasize=15)
    m = ude.NODE(df, hidden_units=4, seed=2)
    sg = ude.SplineGradientMatching(m, T=12, regularization_weight=1e-3)
    x = sg.initial_parameters()
    x[: sg.n_alpha] = 0.5 + 0.3 * np.random.default_rng(0).normal(size=sg.n_alpha)

    class OracleEngine:
        def loss_grad(self, th):
            return oracle.loss_grad(sg.problem, th)

    with ude.Engine(sg.problem) as eng:
        L, g = sg.loss_grad(eng, x)
    Lo, go = sg.loss_grad(OracleEngine(), x)
    d, N, T = 2, 15, 12
    alpha = x[: sg.n_alpha].reshape(sg.n_knots, d, order="F"); pm = x[sg.n_alpha + d * N:]
    t = m.times; dt = (t[-1] - t[0]) / T; Dt = dt / ((t[-1] - t[0]) / N)
    uh = (alpha[sg.i1] * sg.w1[:, None] + alpha[sg.i2] * sg.w2[:, None]).T
    ref = np.sum(((m.data - uh) / 0.05 ** 2) ** 2)
    That = sg.n_knots - 1
    for k in range(That - 1):
        r = (alpha[k + 1] - alpha[k]) / (sg.ts[k + 1] - sg.ts[k]) - onp.rhs(sg.problem, pm, 0, 0, alpha[k], sg.ts[k])
        ref += N / That * np.sum((r / (np.sqrt(Dt) * 0.025 ** 2)) ** 2)
    pr = sg.problem
    ref += 1e-3 * (np.sum(pm[pr.off_w1:pr.off_w1 + 8] ** 2) + np.sum(pm[pr.off_w2:pr.off_w2 + 8] ** 2))
    assert abs(L - ref) < 1e-12 * abs(ref) and abs(L - Lo) < 1e-12 * abs(Lo)
    assert np.max(np.abs(g - go)) < 1e-10 * np.max(np.abs(go))
    ude.train_(m, "spline gradient matching", verbose=False, loss_options=dict(T=12), optim_options=dict(maxiter=30))
    tr = m.extra["loss_trace"]
    assert tr[-1] < tr[0] and m.extra["spline_alpha"].shape == (17, 2) and np.all(np.isfinite(m.uhat))


def test_neural_gradient_matching(ude):
    """train!(model; loss_function = "neural gradient matching") (LFC.jl:764-846; reference test: test/UDETests.jl:43 runs it).
    One evaluation against the formula written out with the NumPy RHS and against the oracle-driven twin, then training."""
    from oracle import ude_oracle_np as onp
    df = ude.LotkaVolterra(plot=False, datasize=10)
    X = pd.DataFrame({"time": np.linspace(-0.5, 4.5, 11), "X": np.sin(np.linspace(0, 3, 11))})
    m = ude.NODE(df, X, hidden_units=4, seed=2)
    ng = ude.NeuralGradientMatching(m, regularization_weight=1e-3, prefit=False)
    x = ng.initial_parameters()
    x = x + 0.01 * np.random.default_rng(0).normal(size=x.shape)

    class OracleEngine:
        def loss_grad(self, th):
            return oracle.loss_grad(ng.problem, th)

    with ude.Engine(ng.problem) as eng:
        L, g = ng.loss_grad(eng, x)
    Lo, go = ng.loss_grad(OracleEngine(), x)
    d, N = 2, 10
    n_theta = m.parameters.shape[0]
    pm = x[d * N:n_theta]
    U, dU, _ = ng.interp(ng.unpack(x[n_theta:]))
    ref = np.sum((m.data - U) ** 2) / 0.1 ** 2
    for j in range(N):
        r = dU[:, j] - onp.rhs(ng.problem, pm, j, j, U[:, j], m.times[j])
        ref += np.sum(r * r) / 0.5 ** 2
    pr = ng.problem
    ref += 1e-3 * (np.sum(pm[pr.off_w1:pr.off_w1 + pr.nn_hidden * pr.nn_in] ** 2) + np.sum(pm[pr.off_w2:pr.off_w2 + pr.nn_out * pr.nn_hidden] ** 2))
    assert abs(L - ref) < 1e-12 * abs(ref) and abs(L - Lo) < 1e-12 * abs(Lo)
    assert np.max(np.abs(g - go)) < 1e-10 * np.max(np.abs(go))
    ude.train_(m, "neural gradient matching", verbose=False, optim_options=dict(maxiter=20))
    tr = m.extra["loss_trace"]
    assert tr[-1] < tr[0] and np.all(np.isfinite(m.uhat)) and np.max(np.abs(m.uhat - m.data)) < 0.5


def test_leave_site_out_and_predict(ude):
    """leave_site_out (src/CrossValidation.jl:463-474) and predict(::MultiUDE, test_data) (src/MultiModelTests.jl:124-150):
    predictions equal the oracle's one-interval flow from the observed start with the covariates of the labelled series."""
    from oracle import ude_oracle_np as onp
    Y, ts = ude.synthetic.lotka_volterra_data(np.random.default_rng(0), 3, 7, random_u0=True)
    dfm = pd.concat([pd.DataFrame({"time": ts, "series": s + 1, "x1": Y[s, 0], "x2": Y[s, 1]}) for s in range(3)], ignore_index=True)
    Xc = pd.concat([pd.DataFrame({"time": np.linspace(-0.5, 4.0, 6), "series": s + 1, "X": np.cos(s + np.linspace(0, 2, 6))}) for s in range(3)],
                   ignore_index=True)
    mm = ude.MultiNODE(dfm, Xc, hidden_units=4, seed=1)
    test = dfm[dfm["series"] == 2].copy(); test["series"] = 3
    out = ude.predict(mm, test)
    assert list(out.columns) == ["series", "time", "x1", "x2"] and len(out) == 6
    prob = ude.default_loss(mm)
    yy = test.sort_values("time")[["x1", "x2"]].to_numpy().T
    for k in range(6):
        ref = np.real(onp.flow(prob, mm.process_model, 2, 2, yy[:, k], ts[k], ts[k + 1]))      # covariates of series index 2 (label 3)
        assert np.max(np.abs(out[["x1", "x2"]].to_numpy()[k] - ref)) < 1e-10
    score = ude.leave_site_out(mm, lambda m: ude.train_(m, "conditional likelihood", verbose=False, optim_options=dict(maxiter=2)))
    assert np.isfinite(score) and score > 0
    # identical predictions from every site => the energy score is the mean absolute error
    p1 = out.copy()
    assert abs(ude.energy_score(test[test["time"] > ts[0]], [p1, p1.copy()], mm)
               - np.mean(np.abs(p1[["x1", "x2"]].to_numpy() - yy[:, 1:].T).sum(axis=1))) < 1e-12
