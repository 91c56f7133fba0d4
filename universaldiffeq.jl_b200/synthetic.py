"""Synthetic workloads of the shapes BASELINE.json names, generated with NumPy.

Restates the reference's simulators as *inputs only* (Julia's Random.seed! streams cannot be
reproduced here, so these are seeded with numpy default_rng):
  LotkaVolterra         /root/reference/src/SimulateData.jl:21-55   (r=4, alpha=5, theta=.75, m=5, u0=(1,.5))
  simulate_coral_data   /root/reference/src/SimulateData.jl:244-276 (coral/algae + decaying covariate)
and builds the flat Problem + an initial theta for configs C1..C5.
"""
from __future__ import annotations

import numpy as np

from . import problem as P


def glorot_f32(rng, out_dim, in_dim):
    """Lux Dense default init: glorot_uniform in Float32, promoted to Float64."""
    lim = np.sqrt(6.0 / (in_dim + out_dim))
    return rng.uniform(-lim, lim, size=(out_dim, in_dim)).astype(np.float32).astype(np.float64)


def init_pm(prob: P.Problem, rng, scalars=(), series_param=None, w_scale: float = 1.0) -> np.ndarray:
    pm = np.zeros(prob.n_pm)
    H, din, dout = prob.nn_hidden, prob.nn_in, prob.nn_out
    pm[prob.off_w1:prob.off_w1 + H * din] = (w_scale * glorot_f32(rng, H, din)).reshape(-1, order="F")
    pm[prob.off_w2:prob.off_w2 + dout * H] = (w_scale * glorot_f32(rng, dout, H)).reshape(-1, order="F")
    for o, v in zip(prob.off_scalar, scalars):
        pm[o] = v
    if series_param is not None:
        pm[prob.off_series_param:prob.off_series_param + len(series_param)] = series_param
    return pm


def _rk4_integrate(f, u0, ts, sub=20):
    """u0: (n, d); ts: (T,) -> (n, d, T) with `sub` RK4 substeps per output interval."""
    n, d = u0.shape
    out = np.empty((n, d, ts.shape[0]))
    u = u0.copy()
    out[:, :, 0] = u
    for k in range(ts.shape[0] - 1):
        h = (ts[k + 1] - ts[k]) / sub
        t = ts[k]
        for _ in range(sub):
            k1 = f(u, t)
            k2 = f(u + 0.5 * h * k1, t + 0.5 * h)
            k3 = f(u + 0.5 * h * k2, t + 0.5 * h)
            k4 = f(u + h * k3, t + h)
            u = u + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
            t += h
        out[:, :, k + 1] = u
    return out


def lotka_volterra_data(rng, n_series=1, n_pts=60, T=3.0, sigma=0.075, random_u0=False, pairs=1):
    """(n_series, 2*pairs, n_pts) noisy LV trajectories and the common time grid."""
    ts = np.linspace(0.0, T, n_pts)
    r, alpha, theta, m = 4.0, 5.0, 0.75, 5.0

    def f(u, t):
        du = np.empty_like(u)
        for q in range(pairs):
            a, b = u[:, 2 * q], u[:, 2 * q + 1]
            du[:, 2 * q] = r * a - alpha * a * b
            du[:, 2 * q + 1] = theta * alpha * a * b - m * b
        return du

    if random_u0:
        u0 = rng.uniform(0.5, 1.5, size=(n_series, 2 * pairs))
    else:
        u0 = np.tile(np.array([1.0, 0.5]), (n_series, pairs))
    clean = _rk4_integrate(f, u0, ts)
    noisy = clean + clean * rng.normal(0.0, sigma, size=clean.shape)
    return noisy, ts


def ou_covariates(rng, n_series, n_knots, T, rho=1.0, sd=0.5):
    kt = np.linspace(0.0, T, n_knots)
    x = np.zeros((n_series, n_knots))
    dt = kt[1] - kt[0]
    a = np.exp(-rho * dt)
    s = sd * np.sqrt(1 - a * a)
    x[:, 0] = rng.normal(0, sd, n_series)
    for k in range(1, n_knots):
        x[:, k] = a * x[:, k - 1] + s * rng.normal(0, 1, n_series)
    return kt, x


def _tile_series(prob: P.Problem, Y, ts):
    """Y: (n_series, d, n_pts) on a common grid -> columns sorted by (series, time)."""
    n, d, T = Y.shape
    prob.data = np.asfortranarray(Y.transpose(1, 0, 2).reshape(d, n * T))
    prob.times = np.tile(ts, n)
    prob.starts = np.arange(n, dtype=np.int64) * T
    prob.lengths = np.full(n, T, dtype=np.int64)


def theta_from(prob: P.Problem, pm_list, uhat_init):
    th = np.zeros(prob.n_theta)
    for m in range(prob.n_models):
        u = prob.uhat_view(th, m)
        c0, c1 = int(prob.model_col0[m]), int(prob.model_col0[m + 1])
        u[:, :] = uhat_init[:, c0:c1] if isinstance(uhat_init, np.ndarray) else uhat_init
        prob.pm_view(th, m)[:] = pm_list[m] if isinstance(pm_list, (list, tuple)) else pm_list
    return th


# ----------------------------------------------------------------------------------------------
def config_c1(seed=123, solver=P.RK4, substeps=4):
    """LV tutorial UDE (README.md:46-65): d=2, NN 2->10->1, (r,k,m,theta)=(.1,10.1,.1,.1), conditional likelihood
    with Sigma = 0.025 I (src/Optimizers.jl:368-369), 60 points => 59 segments."""
    rng = np.random.default_rng(seed)
    Y, ts = lotka_volterra_data(rng, 1, 60)
    prob = P.Problem(d=2, nn_in=2, nn_hidden=10, nn_out=1, rhs_kind=P.RHS_LV_UDE_ABS, n_scalars=4,
                     solver=solver, substeps=substeps, loss_kind=P.LOSS_STATE_SPACE,
                     A=np.eye(2) / 0.025, B=np.eye(2) / 0.025).default_layout()
    _tile_series(prob, Y, ts)
    prob.finalize()
    pm = init_pm(prob, rng, scalars=(0.1, 10.1, 0.1, 0.1))
    theta = theta_from(prob, pm, 1e-3)       # src/helpers.jl:3: uhat = zeros .+ 10^-3
    return prob, theta


def config_c2(n_series=10_000, seed=2, solver=P.RK4, substeps=4, n_pts=60, hidden=10):
    """NODE d=2, hidden 10, multiple shooting (pred_length 5) over n_series synthetic LV series."""
    rng = np.random.default_rng(seed)
    Y, ts = lotka_volterra_data(rng, n_series, n_pts, random_u0=True)
    prob = P.Problem(d=2, nn_in=2, nn_hidden=hidden, nn_out=2, rhs_kind=P.RHS_NODE, solver=solver, substeps=substeps,
                     loss_kind=P.LOSS_MULTI_SHOOT, pred_length=5, preroll_eps=0.0).default_layout()
    _tile_series(prob, Y, ts)
    prob.reg_weight = 1e-6
    prob.finalize()
    pm = init_pm(prob, rng)
    theta = theta_from(prob, pm, np.asarray(prob.data))   # uhat initialised at the data
    return prob, theta


def config_c3(n_series=100_000, seed=3, solver=P.RK4, substeps=4, n_pts=30, hidden=32, loss="multiple shooting"):
    """MultiNODE with covariates: d=4 (two LV pairs), 1 covariate per series (OU path, 30 knots), NN 5->32->4,
    series sharded across GPUs (src/MultiProcessModel.jl:204-242; src/LossFunctionConstructors.jl:639-708)."""
    rng = np.random.default_rng(seed)
    Y, ts = lotka_volterra_data(rng, n_series, n_pts, random_u0=True, pairs=2)
    kt, kx = ou_covariates(rng, n_series, n_pts, ts[-1])
    prob = P.Problem(d=4, n_cov=1, nn_in=5, nn_hidden=hidden, nn_out=4, rhs_kind=P.RHS_NODE, solver=solver,
                     substeps=substeps, pred_length=5).default_layout()
    if loss == "multiple shooting":
        prob.loss_kind = P.LOSS_MULTI_SHOOT
    else:  # MultiNODE default loss: ObservationMSE(T, w) and ProcessMSE(N, T, w) (src/MultipleTimeSeries.jl:189-190)
        prob.loss_kind = P.LOSS_STATE_SPACE
        N, T = n_series * n_pts, float(ts[-1])
        prob.obs_scale = 1.0 / T
        prob.proc_scale = T / (N * N)
        prob.proc_inv_dt = True
    _tile_series(prob, Y, ts)
    prob.knot_off = np.arange(n_series + 1, dtype=np.int64) * n_pts
    prob.knot_t = np.tile(kt, n_series)
    prob.knot_x = kx.reshape(-1)
    prob.reg_weight = 1e-6
    prob.finalize()
    pm = init_pm(prob, rng)
    theta = theta_from(prob, pm, np.asarray(prob.data))
    return prob, theta


def coral_data(rng, n_pts=125):
    """3 states (A, C, X_C) following src/SimulateData.jl:244-276 (Euler, dt = 1/20, every 2nd year)."""
    g, rA, rC, m, b, r, rhoC, mA, bA, rhoA = 0.1, 0.5, 0.3, 3.25, 7.5, 0.0075, 0.50, 5.5, 0.2, 0.4
    T = 2 * n_pts
    u = np.array([0.1, 0.8, 0.0, 0.0])
    out = np.zeros((3, T))
    for t in range(1, T + 1):
        lse = np.log(np.exp(u[0]) + np.exp(u[1]) + np.exp(1 - u[0] - u[1]))
        out[0, t - 1] = np.log(max(u[0], 1e-6)) + lse
        out[1, t - 1] = np.log(max(u[1], 1e-6)) + lse
        out[2, t - 1] = u[2]
        for _ in range(20):
            A, Cc, XC, XA = u
            dA = -A * g / (1 - Cc) - A * np.exp(-mA + bA * XA) + (rA * A + 0.05) * (1 - Cc - A)
            dC = rC * (1 - Cc - A) - np.exp(-m + b * XC + r * t) * Cc
            du = np.array([dA, dC, -rhoC * XC, -rhoA * XA])
            u = u + du / 20 + np.concatenate([[0, 0], rng.normal(0, 1.0, 2)]) / 20
            u[0] = min(max(u[0], 1e-3), 0.98)
            u[1] = min(max(u[1], 1e-3), 0.98)
    return out[:, ::2], np.arange(1, T + 1, 2, dtype=np.float64)


def config_c4(n_members=1000, n_folds=10, seed=4, solver=P.RK4, substeps=4, n_pts=125):
    """Coral UDE (docs/src/examples.md:33-53): NN([u1,u2,u3,t/50-1]) 4->10->2 plus dX = -rho X; an ensemble of
    n_members NN seeds x n_folds leave-future-out folds (fold f trains on the first N-f points,
    src/CrossValidation.jl:23-26) = n_members*n_folds independent theta vectors; conditional likelihood."""
    rng = np.random.default_rng(seed)
    Y, ts = coral_data(rng, n_pts)
    prob = P.Problem(d=3, nn_in=4, nn_hidden=10, nn_out=2, rhs_kind=P.RHS_NODE_TIME, n_scalars=1, solver=solver,
                     substeps=substeps, loss_kind=P.LOSS_STATE_SPACE, time_scale=1.0 / 50.0, time_shift=-1.0,
                     A=np.eye(3) / 0.025, B=np.eye(3) / 0.025).default_layout()
    cols, times, starts, lengths, model = [], [], [], [], []
    c = 0
    for mem in range(n_members):
        for f in range(1, n_folds + 1):
            n = n_pts - f
            cols.append(Y[:, :n]); times.append(ts[:n]); starts.append(c); lengths.append(n)
            model.append(mem * n_folds + (f - 1)); c += n
    prob.data = np.concatenate(cols, axis=1)
    prob.times = np.concatenate(times)
    prob.starts, prob.lengths, prob.model_of_series = np.array(starts), np.array(lengths), np.array(model)
    prob.finalize()
    pms = []
    for mem in range(n_members):
        pm = init_pm(prob, np.random.default_rng(seed * 1000 + mem), scalars=(0.5,))
        pms.extend([pm] * n_folds)             # same seed => same NN init in every fold
    theta = theta_from(prob, pms, np.asarray(prob.data))
    return prob, theta


def config_c5(n_segments=1_000_000, seed=5, solver=P.RK4, substeps=4, hidden=128, d=8):
    """Wide NODE d=8, hidden 128: n_segments independent one-interval segments (conditional-likelihood style)."""
    rng = np.random.default_rng(seed)
    prob = P.Problem(d=d, nn_in=d, nn_hidden=hidden, nn_out=d, rhs_kind=P.RHS_NODE, solver=solver, substeps=substeps,
                     loss_kind=P.LOSS_STATE_SPACE).default_layout()
    Y = rng.normal(0, 1, size=(n_segments, d, 2))
    _tile_series(prob, Y, np.array([0.0, 0.1]))
    prob.finalize()
    pm = init_pm(prob, rng)
    theta = theta_from(prob, pm, np.asarray(prob.data) + 0.05 * rng.normal(size=prob.data.shape))
    return prob, theta


CONFIGS = {"c1": config_c1, "c2": config_c2, "c3": config_c3, "c4": config_c4, "c5": config_c5}
