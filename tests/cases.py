"""Small seeded Problems covering every loss kind x RHS kind x solver the hot path supports, with the edge cases the
reference's data handling produces: ragged series, 1- and 2-point series, trailing single-point multiple-shooting
blocks, irregular time steps, t_skip masks, several independent models (CV folds), covariates with their own knots."""
from __future__ import annotations

import numpy as np

import ude_b200 as ude

P = ude.problem


def _ragged(rng, d, lengths, irregular=True, t0=0.0):
    cols, times, starts = [], [], []
    c = 0
    for n in lengths:
        dt = rng.uniform(0.05, 0.15, size=n) if irregular else np.full(n, 0.1)
        t = t0 + np.cumsum(dt)
        y = 0.8 + 0.4 * rng.normal(size=(d, n)) * 0.5
        cols.append(y); times.append(t); starts.append(c); c += n
    return np.concatenate(cols, axis=1), np.concatenate(times), np.array(starts), np.array(lengths)


def _covariates(rng, prob, n_series, times_by_series, n_knots=7):
    off, kt, kx = [0], [], []
    for s in range(n_series):
        t = times_by_series[s]
        for v in range(prob.n_cov):
            nk = n_knots + v + (s % 3)
            # knots deliberately do not cover the whole span: flat extrapolation on both sides gets exercised
            k = np.sort(rng.uniform(t[0] + 0.02, t[-1] - 0.02, size=nk)) if t[-1] - t[0] > 0.1 else np.array([t[0] - 1.0, t[0] + 1.0])
            kt.append(k); kx.append(rng.normal(size=k.shape[0])); off.append(off[-1] + k.shape[0])
    prob.knot_off = np.array(off); prob.knot_t = np.concatenate(kt); prob.knot_x = np.concatenate(kx)


RHS_SPECS = {
    # name: (rhs_kind, d, n_cov, nn_in, nn_out, n_scalars, has_series_param, scalar init)
    "node_d2": (P.RHS_NODE, 2, 0, 2, 2, 0, False, ()),
    "node_d2x1": (P.RHS_NODE, 2, 1, 3, 2, 0, False, ()),
    "node_d4x1": (P.RHS_NODE, 4, 1, 5, 4, 0, False, ()),
    "node_d8": (P.RHS_NODE, 8, 0, 8, 8, 0, False, ()),
    "node_time_d3": (P.RHS_NODE_TIME, 3, 0, 4, 2, 1, False, (0.5,)),
    "node_time_d2x1": (P.RHS_NODE_TIME, 2, 1, 4, 2, 0, False, ()),
    "lv_ude": (P.RHS_LV_UDE, 2, 0, 2, 1, 3, False, (1.0, 0.5, 0.5)),
    "lv_ude_x1": (P.RHS_LV_UDE, 2, 1, 2, 1, 5, False, (1.0, 0.5, 0.5, 0.1, -0.2)),
    "lv_ude_abs": (P.RHS_LV_UDE_ABS, 2, 0, 2, 1, 4, False, (0.1, 10.1, -0.1, 0.1)),
    "fishery": (P.RHS_FISHERY, 2, 1, 3, 1, 3, False, (1.0, 0.5, 4.0)),
    "nn_decay_d1": (P.RHS_NN_DECAY, 1, 0, 1, 1, 1, False, (0.5,)),
    "series_growth_d1": (P.RHS_SERIES_GROWTH, 1, 0, 1, 1, 0, True, ()),
    "node_d1": (P.RHS_NODE, 1, 0, 1, 1, 0, False, ()),
    "node_d3": (P.RHS_NODE, 3, 0, 3, 3, 0, False, ()),
    "node_d4": (P.RHS_NODE, 4, 0, 4, 4, 0, False, ()),
    "node_d1x1": (P.RHS_NODE, 1, 1, 2, 1, 0, False, ()),
    "node_d3x1": (P.RHS_NODE, 3, 1, 4, 3, 0, False, ()),
    "node_d2x2": (P.RHS_NODE, 2, 2, 4, 2, 0, False, ()),
}
LOSSES = ("cond_lik", "mse", "multi_shoot", "multi_shoot_preroll", "shoot_theta", "shoot_data", "deriv_match")


def make_case(rhs="node_d2", loss="cond_lik", solver=P.RK4, hidden=10, lengths=(7, 1, 12, 2, 6), n_models=1, seed=0,
              substeps=3, skip=False, reg=1e-3, pred_length=5):
    rng = np.random.default_rng(seed)
    kind, d, n_cov, nn_in, nn_out, n_sc, has_sp, sc0 = RHS_SPECS[rhs]
    lengths = list(lengths)
    if loss.startswith("shoot") or loss == "deriv_match":
        lengths = [max(n, 3) for n in lengths]
    prob = P.Problem(d=d, nn_in=nn_in, nn_hidden=hidden, nn_out=nn_out, rhs_kind=kind, n_cov=n_cov, n_scalars=n_sc,
                     has_series_param=has_sp, solver=solver, substeps=substeps, time_scale=0.5, time_shift=-0.25)
    n_series = len(lengths)
    if n_models > 1:
        assert n_series % n_models == 0
        prob.model_of_series = np.repeat(np.arange(n_models), n_series // n_models)
        spm = n_series // n_models
    else:
        spm = n_series
    prob.default_layout(n_series_param=spm)
    Y, t, starts, lens = _ragged(rng, d, lengths)
    prob.data, prob.times, prob.starts, prob.lengths = Y, t, starts, lens
    if n_cov:
        _covariates(rng, prob, n_series, [t[s:s + n] for s, n in zip(starts, lens)])
    M = n_models
    if loss == "cond_lik":
        prob.loss_kind = P.LOSS_STATE_SPACE
        Ao = rng.normal(size=(d, d)); Bo = rng.normal(size=(d, d))
        prob.A = np.linalg.inv(Ao @ Ao.T + d * np.eye(d)) * 40.0            # full Sigma^-1 (errors_to_matrix matrix branch)
        prob.B = np.linalg.inv(np.diag(rng.uniform(0.02, 0.05, size=d)))      # vector branch
    elif loss == "mse":
        prob.loss_kind = P.LOSS_STATE_SPACE
        prob.proc_inv_dt = True
        prob.obs_scale = rng.uniform(0.5, 2.0, size=M)
        prob.proc_scale = rng.uniform(0.5, 2.0, size=M)
    elif loss in ("multi_shoot", "multi_shoot_preroll"):
        prob.loss_kind = P.LOSS_MULTI_SHOOT
        prob.pred_length = pred_length
        prob.preroll_eps = 1e-6 if loss.endswith("preroll") else 0.0
    elif loss in ("shoot_theta", "shoot_data"):
        prob.loss_kind = P.LOSS_SHOOT
        prob.shoot_from_theta = loss == "shoot_theta"
        prob.shoot_weight = rng.uniform(0.5, 2.0, size=M)
    elif loss == "deriv_match":
        prob.loss_kind = P.LOSS_DERIV_MATCH
        prob.remove_ends = 1
        prob.dm_u = Y + 0.01 * rng.normal(size=Y.shape)
        prob.dm_dudt = rng.normal(size=Y.shape)
    else:
        raise ValueError(loss)
    if skip:
        m = np.zeros(Y.shape[1], dtype=np.uint8)
        for s, n in zip(starts, lens):
            if n >= 3:
                m[s + 1 + int(rng.integers(0, n - 1))] = 1
        prob.skip_mask = m
    prob.reg_weight = reg * (1.0 + np.arange(M))
    prob.finalize()
    pms = [ude.synthetic.init_pm(prob, np.random.default_rng(seed * 100 + m), scalars=sc0,
                                 series_param=(rng.uniform(0.2, 1.0, size=spm) if has_sp else None)) for m in range(M)]
    theta = ude.synthetic.theta_from(prob, pms, np.asarray(prob.data) + 0.05 * rng.normal(size=prob.data.shape))
    # biases are zero at Lux init; perturb everything so no gradient entry is trivially zero
    theta = theta + 0.02 * rng.normal(size=theta.shape)
    return prob, theta


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def leaf_groups(prob, m=0):
    """(name, lo, hi) index ranges of model m's gradient leaves inside theta: uhat, W1, b1, W2, b2, each scalar, the
    per-series vector (SURVEY.md App. B layout)."""
    b0 = int(prob.theta_base[m]); pb = prob.pm_base(m)
    H, din, dout = prob.nn_hidden, prob.nn_in, prob.nn_out
    g = [("uhat", b0, pb), ("W1", pb + prob.off_w1, pb + prob.off_w1 + H * din), ("b1", pb + prob.off_b1, pb + prob.off_b1 + H),
         ("W2", pb + prob.off_w2, pb + prob.off_w2 + dout * H), ("b2", pb + prob.off_b2, pb + prob.off_b2 + dout)]
    for k, o in enumerate(prob.off_scalar):
        g.append((f"scalar{k}", pb + o, pb + o + 1))
    if prob.has_series_param:
        g.append(("series_param", pb + prob.off_series_param, pb + prob.n_pm))
    return g


def grad_mismatch(prob, g, g_ref, rtol=1e-10, floor=1e-3):
    """Element-wise gradient check per leaf group: every entry must satisfy
    |g - g_ref| <= rtol |g_ref| + rtol * floor * max|g_ref over its leaf group|.  A norm-wise check over the whole vector
    lets a small-magnitude leaf (b2, scalar parameters) be 100% wrong; this does not.  Returns the worst violation as
    (ratio, model, leaf) with ratio <= 1 meaning pass."""
    worst = (0.0, -1, "")
    for m in range(prob.n_models):
        for name, lo, hi in leaf_groups(prob, m):
            if hi <= lo:
                continue
            a = np.asarray(g[lo:hi], dtype=np.float64); b = np.asarray(g_ref[lo:hi], dtype=np.float64)
            tol = rtol * np.abs(b) + rtol * floor * max(float(np.max(np.abs(b))), 1e-300)
            r = float(np.max(np.abs(a - b) / tol))
            if not np.isfinite(r):
                r = np.inf
            if r > worst[0]:
                worst = (r, m, name)
    return worst

