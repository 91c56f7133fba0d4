"""NumPy twin of the CPU oracle: LOSS ONLY, dtype-generic.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (no Julia in this image; see oracle/ude_oracle.h).  This file exists so that the
hand-derived discrete adjoint in ude_oracle.c can be checked against complex-step derivatives
(theta + i*1e-30*e_k) of an independently written forward evaluation of the same losses:
  /root/reference/src/LossFunctionConstructors.jl:17-98, :175-305, :347-404, :500-708
  /root/reference/src/helpers.jl:16-76, src/MultipleTimeSeries.jl:51-157, src/LossFunctions.jl:17-28
Plain Python loops: small cases only.
"""
from __future__ import annotations

import numpy as np

RK4_TAB = dict(
    c=[0.0, 0.5, 0.5, 1.0],
    a=[[], [0.5], [0.0, 0.5], [0.0, 0.0, 1.0]],
    b=[1 / 6, 1 / 3, 1 / 3, 1 / 6],
)
EULER_TAB = dict(c=[0.0], a=[[]], b=[1.0])
TSIT5_TAB = dict(
    c=[0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0],
    a=[[],
       [0.161],
       [-0.008480655492356989, 0.335480655492357],
       [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
       [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
       [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]],
    b=[0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
)


def _abs_cs(z):
    """abs() that stays analytic for complex-step: z * sign(Re z) (sign(0) := +1, as in the C oracle)."""
    return z * (-1.0 if np.real(z) < 0 else 1.0)


def covariates_at(p, series, t):
    out = np.zeros(p.n_cov)
    for v in range(p.n_cov):
        k0, k1 = int(p.knot_off[series * p.n_cov + v]), int(p.knot_off[series * p.n_cov + v + 1])
        kt, kx = p.knot_t[k0:k1], p.knot_x[k0:k1]
        if t <= kt[0]:
            out[v] = kx[0]
        elif t >= kt[-1]:
            out[v] = kx[-1]
        else:
            lo = int(np.searchsorted(kt, t, side="right") - 1)
            f = (t - kt[lo]) / (kt[lo + 1] - kt[lo])
            out[v] = (1.0 - f) * kx[lo] + f * kx[lo + 1]
    return out


def rhs(p, pm, series_local, series_global, u, t):
    d, H, din, dout = p.d, p.nn_hidden, p.nn_in, p.nn_out
    W1 = pm[p.off_w1:p.off_w1 + H * din].reshape(H, din, order="F")
    b1 = pm[p.off_b1:p.off_b1 + H]
    W2 = pm[p.off_w2:p.off_w2 + dout * H].reshape(dout, H, order="F")
    b2 = pm[p.off_b2:p.off_b2 + dout]
    X = covariates_at(p, series_global, t) if p.n_cov else np.zeros(0)
    k = p.rhs_kind
    if k >= 1000:      # expression right-hand side (universaldiffeq.jl_b200/user_rhs.py): NN([u; X (; a t + b)]), then the user's du
        ur = p.user_rhs
        x = np.concatenate([u, X, [p.time_scale * t + p.time_shift]]) if ur.has_time_input else np.concatenate([u, X])
        y = W2 @ np.tanh(W1 @ x + b1) + b2
        return ur.du_numpy(u, X, y, [pm[o] for o in p.off_scalar], pm[p.off_series_param + series_local] if p.has_series_param else None)
    if k in (0, 4):
        x = np.concatenate([u, X])
    elif k == 1:
        x = np.concatenate([u, X, [p.time_scale * t + p.time_shift]])
    else:
        x = u
    y = W2 @ np.tanh(W1 @ x + b1) + b2
    sc = [pm[o] for o in p.off_scalar]
    if k == 0:
        return y
    if k == 1:
        return np.array([y[j] if j < dout else -sc[j - dout] * u[j] for j in range(d)])
    if k == 2:
        r, m, th = sc[0], sc[1], sc[2]
        du = [r * u[0] - y[0], th * y[0] - m * u[1]]
        if p.n_scalars >= 5:
            du[0] = du[0] + sc[3] * X[0]
            du[1] = du[1] + sc[4] * X[0]
        return np.array(du)
    if k == 3:
        r, kk, m, th = (_abs_cs(s) for s in sc[:4])
        Cc = _abs_cs(y[0])
        return np.array([r * u[0] * (1.0 - u[0] / kk) - Cc, th * Cc - m * u[1]])
    if k == 4:
        q, r, K = sc[0], sc[1], sc[2]
        return np.array([y[0], r * u[1] * (1.0 - u[1] / K) - q * u[0]])
    if k == 5:
        return y - sc[0] * u
    if k == 6:
        return pm[p.off_series_param + series_local] * u * y[0]
    raise ValueError(k)


def rk_step(p, pm, sl, sg, u, t, h):
    tab = TSIT5_TAB if p.solver == 1 else (EULER_TAB if p.solver == 2 else RK4_TAB)
    ks = []
    for i in range(len(tab["c"])):
        U = u
        if i > 0:
            acc = 0
            for l in range(i):
                acc = acc + tab["a"][i][l] * ks[l]
            U = u + h * acc
        ks.append(rhs(p, pm, sl, sg, U, t + tab["c"][i] * h))
    acc = 0
    for i in range(len(ks)):
        acc = acc + tab["b"][i] * ks[i]
    return u + h * acc


def flow(p, pm, sl, sg, u, t0, t1):
    h = (t1 - t0) / p.substeps
    for n in range(p.substeps):
        u = rk_step(p, pm, sl, sg, u, t0 + n * h, h)
    return u


def loss(p, theta):
    """loss[n_models]; theta may be complex."""
    theta = np.asarray(theta)
    d = p.d
    out = np.zeros(p.n_models, dtype=theta.dtype)
    for s in range(p.n_series):
        m = int(p.model_of_series[s])
        ncm = int(p.model_col0[m + 1] - p.model_col0[m])
        base = int(p.theta_base[m])
        uh = theta[base:base + d * ncm].reshape(d, ncm, order="F")
        pm = theta[base + d * ncm: base + d * ncm + p.n_pm]
        c0 = int(p.starts[s])
        n = int(p.lengths[s])
        lc0 = c0 - int(p.model_col0[m])
        sl = s - int(p.model_series0[m])
        y = p.data[:, c0:c0 + n]
        tt = p.times[c0:c0 + n]
        L = 0
        if p.loss_kind == 0:
            for t in range(n):
                e = y[:, t] - uh[:, lc0 + t]
                L = L + p.obs_scale[m] * (e @ (p.A @ e))
            for t in range(1, n):
                if p.skip_mask is not None and p.skip_mask[c0 + t]:
                    continue
                dt = tt[t] - tt[t - 1]
                r = uh[:, lc0 + t] - flow(p, pm, sl, s, uh[:, lc0 + t - 1], tt[t - 1], tt[t])
                w = p.proc_scale[m] * ((1.0 / dt) if p.proc_inv_dt else 1.0)
                L = L + w * (r @ (p.B @ r))
        elif p.loss_kind == 1:
            for t in range(0, n, p.pred_length):
                Lb = p.pred_length if (n - 1 - t) >= p.pred_length else (n - 1 - t)
                u = uh[:, lc0 + t]
                if p.preroll_eps > 0:
                    u = rk_step(p, pm, sl, s, u, tt[t] - p.preroll_eps, p.preroll_eps)
                for i in range(Lb + 1):
                    r = y[:, t + i] - u
                    L = L + (r @ r) / (d * n)
                    if i < Lb:
                        u = flow(p, pm, sl, s, u, tt[t + i], tt[t + i + 1])
        elif p.loss_kind == 2:
            u = uh[:, lc0] if p.shoot_from_theta else y[:, 0].astype(theta.dtype)
            for i in range(n):
                if i >= p.shoot_first_scored:
                    r = y[:, i] - u
                    L = L + p.shoot_weight[m] * (r @ r)
                if i < n - 1:
                    u = flow(p, pm, sl, s, u, tt[i], tt[i + 1])
        else:
            for t in range(p.remove_ends, n - p.remove_ends):
                f = rhs(p, pm, sl, s, p.dm_u[:, c0 + t].astype(theta.dtype), tt[t])
                r = p.dm_dudt[:, c0 + t] - f
                L = L + r @ r
        out[m] = out[m] + L
    for m in range(p.n_models):
        if p.reg_weight[m] != 0:
            pm = theta[p.pm_base(m): p.pm_base(m) + p.n_pm]
            W1 = pm[p.off_w1:p.off_w1 + p.nn_hidden * p.nn_in]
            W2 = pm[p.off_w2:p.off_w2 + p.nn_out * p.nn_hidden]
            out[m] = out[m] + p.reg_weight[m] * (np.sum(W1 * W1) + np.sum(W2 * W2))
    return out


def complex_step_grad(p, theta, model=0, idx=None, eps=1e-30):
    theta = np.asarray(theta, dtype=np.float64)
    idx = range(theta.shape[0]) if idx is None else idx
    g = {}
    for k in idx:
        th = theta.astype(np.complex128)
        th[k] += 1j * eps
        g[k] = float(np.imag(loss(p, th)[model]) / eps)
    return g
