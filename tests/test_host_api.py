"""Host-side mirror of the reference's constructors and loss builders: the Problems they emit encode the reference's
formulas, checked by evaluating them with the oracle against a from-scratch NumPy evaluation."""
import numpy as np
import pandas as pd
import pytest

from oracle import oracle, ude_oracle_np as onp


def _lv_frame(ude, n=25, seed=3):
    Y, ts = ude.synthetic.lotka_volterra_data(np.random.default_rng(seed), 1, n)
    return pd.DataFrame({"time": ts, "x1": Y[0, 0], "x2": Y[0, 1]})


def _readme_model(ude, df):
    _, p = ude.SimpleNeuralNetwork(2, 1)
    return ude.CustomDerivatives(df, ude.LV_UDE_ABS, {"NN": p, "r": 0.1, "k": 10.1, "m": 0.1, "theta": 0.1})


def test_constructor_layout_and_defaults(ude):
    df = _lv_frame(ude).sample(frac=1.0, random_state=0)          # process_data sorts by time (src/helpers.jl:495)
    m = _readme_model(ude, df)
    assert np.all(np.diff(m.times) > 0) and m.data.shape == (2, 25)
    assert m.parameters.shape[0] == 2 * 25 + 41 + 4                # uhat | NN (2->10->1 = 41) | r k m theta
    assert np.all(m.uhat == 1e-3)                                  # src/helpers.jl:3
    assert np.allclose(m.process_model[-4:], [0.1, 10.1, 0.1, 0.1])
    assert m.weights == dict(regularization=1e-6, process=1.0, observation=1.0)
    m2 = ude.MultiNODE(pd.concat([df.assign(series="b"), df.assign(series="a")]), hidden_units=7)
    assert np.all(m2.uhat == 0.0) and m2.obs_divisor == m2.T       # src/MultipleTimeSeries.jl:200, :190
    assert list(m2.starts) == [0, 25] and m2.rhs.nn_in == 2 and m2.hidden == 7


def test_conditional_likelihood_formula(ude):
    df = _lv_frame(ude)
    m = _readme_model(ude, df)
    m.parameters[:50] = m.data.reshape(-1, order="F") + 0.01
    prob = ude.conditional_likelihood(m, regularization_weight=2.0, observation_error=0.05, process_error=[0.02, 0.04])
    L, _ = oracle.loss_grad(prob, m.parameters)
    # App. A.1 by hand
    uh, pm = m.uhat, m.process_model
    So, Sp = np.linalg.inv(0.05 * np.eye(2)), np.linalg.inv(np.diag([0.02, 0.04]))
    ref = sum((m.data[:, t] - uh[:, t]) @ So @ (m.data[:, t] - uh[:, t]) for t in range(25))
    for t in range(1, 25):
        r = uh[:, t] - np.real(onp.flow(prob, pm, 0, 0, uh[:, t - 1], m.times[t - 1], m.times[t]))
        ref += r @ Sp @ r
    W1 = pm[:20]; W2 = pm[30:40]
    ref += 2.0 * 1e-6 * (np.sum(W1 ** 2) + np.sum(W2 ** 2))       # train!'s weight times the constructor's (App. A.1)
    assert abs(L[0] - ref) < 1e-11 * abs(ref)


def test_default_loss_and_multi_reg_quirk(ude):
    df = _lv_frame(ude)
    dfm = pd.concat([df.assign(series=1), df.iloc[:14].assign(series=2)])
    m = ude.MultiNODE(dfm, reg_weight=1e-3, obs_weight=0.5, proc_weight=2.0)
    m.parameters[: m.d * m.N] = m.data.reshape(-1, order="F") * 0.9
    prob = ude.default_loss(m)
    L, _ = oracle.loss_grad(prob, m.parameters)
    uh, pm = m.uhat, m.process_model
    ref = 0.0
    for s, (c0, n) in enumerate(zip(m.starts, m.lengths)):
        for t in range(n):
            ref += 0.5 * np.sum((m.data[:, c0 + t] - uh[:, c0 + t]) ** 2) / m.T                    # ObservationMSE(T, w)
        for t in range(1, n):
            dt = m.times[c0 + t] - m.times[c0 + t - 1]
            r = uh[:, c0 + t] - np.real(onp.flow(prob, pm, s, s, uh[:, c0 + t - 1], m.times[c0 + t - 1], m.times[c0 + t]))
            ref += m.T / dt * 2.0 * np.sum(r ** 2) / m.N ** 2                                        # ProcessMSE(N, T, w)
        ref += 1e-3 * (np.sum(pm[prob.off_w1:prob.off_w1 + 20] ** 2) + np.sum(pm[prob.off_w2:prob.off_w2 + 20] ** 2))  # once PER SERIES
    assert abs(L[0] - ref) < 1e-11 * abs(ref)


def test_t_skip_rules_differ_between_losses(ude):
    """conditional_likelihood skips the interval whose 1-based END INDEX equals t_skip (LFC.jl:82); the default loss
    UDE.loss_function(x, t_skip) skips intervals whose START TIME VALUE is in t_skip (src/helpers.jl:63)."""
    df = _lv_frame(ude, n=12)
    m = ude.NODE(df, hidden_units=4)
    a = ude.conditional_likelihood(m, t_skip=5).skip_mask
    assert a.sum() == 1 and a[4] == 1                                   # interval 4 -> 5 (1-based), stored at its end column
    b = ude.default_loss(m, t_skip=[m.times[2], m.times[7], 123.0]).skip_mask
    assert b.sum() == 2 and b[3] == 1 and b[8] == 1                     # intervals STARTING at times[2] and times[7]
    assert ude.default_loss(m, t_skip=5).skip_mask.sum() == int(np.any(m.times[:-1] == 5.0))


def test_shooting_and_multiple_shooting_builders(ude):
    df = _lv_frame(ude, n=16)
    m = ude.NODE(df, hidden_units=5)
    ms = ude.multiple_shooting_loss(m, 0.0, pred_length=5)
    assert ms.preroll_eps == 1e-6 and ms.count_segments() == 4 and ms.count_steps() == 15 * 4 + 4   # LFC.jl:268 pre-roll
    sh = ude.shooting_loss(m)
    assert sh.shoot_from_theta and sh.shoot_first_scored == 2
    assert np.isclose(sh.shoot_weight[0], m.T * 1.0 / 16 ** 2) and np.isclose(sh.reg_weight[0], 1e-6)
    mm = ude.MultiNODE(pd.concat([df.assign(series=1), df.assign(series=2)]))
    assert ude.multiple_shooting_loss(mm).preroll_eps == 0.0 and not ude.shooting_loss(mm).shoot_from_theta
    with pytest.raises(ValueError):
        ude.build_problem(m, "marginal likelihood")


def test_covariates_long_and_wide(ude):
    df = _lv_frame(ude, n=12)
    Xw = pd.DataFrame({"time": np.linspace(0, 3, 7), "X1": np.arange(7.0), "X2": -np.arange(7.0)})
    m = ude.NODE(df, Xw, hidden_units=4)
    assert m.n_cov == 2 and m.rhs.nn_in == 4
    Xl = Xw.melt(id_vars="time", var_name="variable", value_name="value")
    m2 = ude.NODE(df, Xl, variable_column_name="variable", value_column_name="value", hidden_units=4)
    assert np.array_equal(m.cov[2], m2.cov[2]) and np.array_equal(m.cov[3], m2.cov[3])
    p = ude.conditional_likelihood(m)
    assert np.allclose(onp.covariates_at(p, 0, 0.25), [0.5, -0.5]) and np.allclose(onp.covariates_at(p, 0, 99.0), [6, -6])


def test_errors_to_matrix(ude):
    assert np.array_equal(ude.errors_to_matrix(0.5, 2), 0.5 * np.eye(2))
    assert np.array_equal(ude.errors_to_matrix([1, 2], 2), np.diag([1.0, 2.0]))
    M = np.array([[1.0, 0.2], [0.2, 2.0]])
    assert ude.errors_to_matrix(M, 2) is not None and np.array_equal(ude.errors_to_matrix(M, 2), M)
    with pytest.raises(ValueError):
        ude.errors_to_matrix([1, 2, 3], 2)


def test_stack_models_is_independent_folds(ude):
    df = _lv_frame(ude, n=14)
    m = _readme_model(ude, df)
    folds = [m.constructor(df.iloc[:-i]) for i in (1, 2, 3)]
    for i, f in enumerate(folds):
        f.parameters[: 2 * f.N] = f.data.reshape(-1, order="F") + 0.01 * (i + 1)
    big, theta = ude.stack_models(folds, "conditional likelihood")
    L, g = oracle.loss_grad(big, theta)
    for i, f in enumerate(folds):
        Li, gi = oracle.loss_grad(ude.conditional_likelihood(f), f.parameters)
        lo = int(big.theta_base[i])
        assert abs(L[i] - Li[0]) < 1e-12 * abs(Li[0]) and np.allclose(g[lo:lo + gi.shape[0]], gi, rtol=1e-12, atol=1e-14)


def test_simulated_data_sets(ude):
    """LotkaVolterra / LorenzLotkaVolterra / simulate_coral_data (src/SimulateData.jl:21-55, :80-118, :244-276): shapes, columns,
    determinism in the seed, and the noise-free limit against SciPy's adaptive solver."""
    from scipy.integrate import solve_ivp
    df = ude.LotkaVolterra(plot=False, seed=7, datasize=40, T=2.0, sigma=0.0)
    assert list(df.columns) == ["time", "x1", "x2"] and len(df) == 40
    sol = solve_ivp(lambda t, u: [4 * u[0] - 5 * u[0] * u[1], 0.75 * 5 * u[0] * u[1] - 5 * u[1]], (0, 2.0), [1.0, 0.5],
                    t_eval=df["time"].to_numpy(), rtol=1e-10, atol=1e-12)
    assert np.max(np.abs(sol.y[0] - df["x1"])) < 1e-5 and np.max(np.abs(sol.y[1] - df["x2"])) < 1e-5
    a, b = ude.LotkaVolterra(seed=1), ude.LotkaVolterra(seed=1)
    assert a.equals(b) and not a.equals(ude.LotkaVolterra(seed=2))
    data, X = ude.LorenzLotkaVolterra(plot=False, datasize=30)
    assert list(data.columns) == ["time", "x1", "x2"] and list(X.columns) == ["time", "X"] and len(X) == 30
    assert np.all(np.isfinite(data.to_numpy())) and np.all(np.isfinite(X.to_numpy()))
    data, X = ude.simulate_coral_data()
    assert list(data.columns) == ["time", "A", "C"] and len(data) == 125 and data["time"].iloc[-1] == 249


def test_save_and_load_model_parameters(ude, tmp_path):
    """save_model_parameters / load_model_parameters! (src/LoadSaveModels.jl:1-11)."""
    df = ude.LotkaVolterra(plot=False, datasize=12)
    a = ude.NODE(df, hidden_units=5, seed=1)
    b = ude.NODE(df, hidden_units=5, seed=2)
    assert not np.array_equal(a.process_model, b.process_model)
    f = tmp_path / "pm.json"
    ude.save_model_parameters(a, f)
    ude.load_model_parameters_(b, f)
    assert np.array_equal(a.process_model, b.process_model)
    c = ude.NODE(df, hidden_units=6, seed=2)
    with pytest.raises(ValueError):
        ude.load_model_parameters_(c, f)


class _OracleEngine:
    """stands in for ude.Engine on CPU: the C oracle evaluates the same Problem (test infrastructure only)."""

    def __init__(self, prob):
        self.prob = prob

    def loss_grad(self, th):
        from oracle import oracle
        return oracle.loss_grad(self.prob, th)


def test_spline_gradient_matching_formula_on_cpu(ude):
    """the Euler / state-space restatement of spline_gradient_matching_loss (LFC.jl:888-931) equals the formula written out
    with the NumPy RHS; its gradient equals central differences."""
    from oracle import ude_oracle_np as onp
    df = ude.LotkaVolterra(plot=False, datasize=15)
    m = ude.NODE(df, hidden_units=4, seed=2)
    sg = ude.SplineGradientMatching(m, T=12, regularization_weight=1e-3)
    assert sg.n_knots == 17 and sg.problem.solver == ude.problem.EULER
    x = sg.initial_parameters()
    x[: sg.n_alpha] = 0.5 + 0.3 * np.random.default_rng(0).normal(size=sg.n_alpha)
    eng = _OracleEngine(sg.problem)
    L, g = sg.loss_grad(eng, x)
    d, N, T = 2, 15, 12
    alpha = x[: sg.n_alpha].reshape(sg.n_knots, d, order="F"); pm = x[sg.n_alpha + d * N:]
    t = m.times; dt = (t[-1] - t[0]) / T; Dt = dt / ((t[-1] - t[0]) / N)
    # the reference's interpolation: weight (t - ts1)/(ts2 - ts1) on the LEFT knot
    uh = np.stack([alpha[sg.i1[j]] * sg.w1[j] + alpha[sg.i2[j]] * sg.w2[j] for j in range(N)], axis=1)
    ref = np.sum(((m.data - uh) / 0.05 ** 2) ** 2)
    That = sg.n_knots - 1
    for k in range(That - 1):
        r = (alpha[k + 1] - alpha[k]) / (sg.ts[k + 1] - sg.ts[k]) - onp.rhs(sg.problem, pm, 0, 0, alpha[k], sg.ts[k])
        ref += N / That * np.sum((r / (np.sqrt(Dt) * 0.025 ** 2)) ** 2)
    pr = sg.problem
    ref += 1e-3 * (np.sum(pm[pr.off_w1:pr.off_w1 + 8] ** 2) + np.sum(pm[pr.off_w2:pr.off_w2 + 8] ** 2))
    assert abs(L - ref) < 1e-12 * abs(ref)
    for k in (0, 5, sg.n_alpha - 1, sg.n_alpha + d * N + 3):
        e = np.zeros_like(x); h = 1e-6 * max(1.0, abs(x[k])); e[k] = h
        fd = (sg.loss_grad(eng, x + e)[0] - sg.loss_grad(eng, x - e)[0]) / (2 * h)
        assert abs(g[k] - fd) <= 1e-6 * max(1.0, abs(fd))


def test_neural_gradient_matching_formula_on_cpu(ude):
    """neural_gradient_matching_loss (LFC.jl:764-846) through Euler mini-series, with covariates, against the formula and
    central differences; the pre-fit of the interpolating network reproduces the data."""
    from oracle import ude_oracle_np as onp
    df = ude.LotkaVolterra(plot=False, datasize=10)
    X = pd.DataFrame({"time": np.linspace(-0.5, 4.5, 11), "X": np.sin(np.linspace(0, 3, 11))})
    m = ude.NODE(df, X, hidden_units=4, seed=2)
    ng = ude.NeuralGradientMatching(m, regularization_weight=1e-3, prefit=False)
    x = ng.initial_parameters()
    x = x + 0.01 * np.random.default_rng(0).normal(size=x.shape)
    eng = _OracleEngine(ng.problem)
    L, g = ng.loss_grad(eng, x)
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
    assert abs(L - ref) < 1e-12 * abs(ref)
    for k in (d * N + 2, n_theta + 1, n_theta + N + 3, n_theta + 2 * N + 5, len(x) - 1):
        e = np.zeros_like(x); e[k] = 1e-6
        fd = (ng.loss_grad(eng, x + e)[0] - ng.loss_grad(eng, x - e)[0]) / 2e-6
        assert abs(g[k] - fd) <= 1e-6 * max(1.0, abs(fd))
    fit = ude.NeuralGradientMatching(m, prefit=True)
    assert np.sum((m.data - fit.interp(fit.ip)[0]) ** 2) < 1e-3
