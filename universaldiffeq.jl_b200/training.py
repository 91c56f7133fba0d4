"""Host-side mirror of the reference's loss constructors and training driver for the hot path.

  errors_to_matrix, conditional_likelihood   /root/reference/src/LossFunctionConstructors.jl:4-98, :347-404
  derivative_matching_loss                   :161-192, :500-540
  shooting_loss / multiple_shooting_loss     :196-305, :545-708        (+ *_states :237-258, :310-341, :604-631, :711-756)
  default loss (init_loss)                   src/helpers.jl:16-76, src/MultipleTimeSeries.jl:51-157
  train!, gradient_descent!, default_options src/Optimizers.jl:12-39, :242-258, :352-570, :578-742
Each constructor returns a finalized Problem; the loss closure + Zygote pullback of the reference is
Engine(problem).loss_grad.  Adam follows Optimisers.jl.
"""
from __future__ import annotations

import copy
from typing import Optional

import numpy as np

from . import problem as P
from . import smoothing
from ._capi import Engine
from .models import UDE


def errors_to_matrix(error, dims):
    """src/LossFunctionConstructors.jl:4-15: scalar -> e*I, vector -> diag(e), matrix -> itself."""
    e = np.asarray(error, dtype=np.float64)
    if e.ndim == 0:
        return float(e) * np.eye(dims)
    if e.ndim == 1 and e.shape[0] == dims:
        return np.diag(e)
    if e.shape == (dims, dims):
        return e
    raise ValueError("error term needs to be a float, vector or square matrix")


def _n_series(model: UDE) -> int:
    return len(model.starts) if model.multi else 1


def _skip_mask(model: UDE, t_skip, by="index"):
    """Column mask of the process terms left out.  The reference has TWO rules for single-series models:
      by="index"  conditional_likelihood: `if t != t_skip` on the 1-based END index of the interval (LFC.jl:82);
      by="time"   UDE.loss_function(x, t_skip) -- the default loss: `if !(times[t-1] in t_skip)`, i.e. intervals whose START
                  TIME VALUE is in t_skip, a scalar or a collection (src/helpers.jl:63).
    Multi-series models match start times per series in both (src/MultipleTimeSeries.jl:101)."""
    if t_skip is None or (np.isscalar(t_skip) and np.isnan(t_skip)):
        return None
    m = np.zeros(model.data.shape[1], dtype=np.uint8)
    if not model.multi and by == "time":
        tt = np.asarray(model.times)
        for ts in np.atleast_1d(np.asarray(t_skip, dtype=np.float64)):
            for j in np.nonzero(tt[:-1] == ts)[0]:
                m[j + 1] = 1
    elif not model.multi:
        # LFC.jl:82: the interval whose END index t (1-based) equals t_skip is left out
        idx = int(t_skip) - 1
        if 1 <= idx < m.shape[0]:
            m[idx] = 1
    else:
        # src/MultipleTimeSeries.jl:101: intervals STARTING at time t_skip (per series) are left out
        for s, n in zip(model.starts, model.lengths):
            tt = model.times[s:s + n]
            for j in np.nonzero(np.isclose(tt[:-1], t_skip))[0]:
                m[s + j + 1] = 1
    return m


def conditional_likelihood(model: UDE, regularization_weight=0.0, observation_error=0.025, process_error=0.025, t_skip=None) -> P.Problem:
    p = model.base_problem()
    p.loss_kind = P.LOSS_STATE_SPACE
    p.A = np.linalg.inv(errors_to_matrix(observation_error, model.d))
    p.B = np.linalg.inv(errors_to_matrix(process_error, model.d))
    p.skip_mask = _skip_mask(model, t_skip)
    p.reg_weight = regularization_weight * model.weights["regularization"]
    return p.finalize()


def default_loss(model: UDE, t_skip=None) -> P.Problem:
    """model.loss_function of the reference: sum_t w_o |y-uhat|^2/N + sum_{t>=2} (T/dt) w_p |uhat_t - F(uhat_{t-1})|^2/N^2 + reg."""
    p = model.base_problem()
    p.loss_kind = P.LOSS_STATE_SPACE
    p.proc_inv_dt = True
    p.obs_scale = model.weights["observation"] / model.obs_divisor
    p.proc_scale = model.T * model.weights["process"] / float(model.N) ** 2
    p.skip_mask = _skip_mask(model, t_skip, by="time")
    # multi-series: the regulariser sits inside single_loss and is counted once per series (src/MultipleTimeSeries.jl:77-79)
    p.reg_weight = model.weights["regularization"] * _n_series(model)
    return p.finalize()


def multiple_shooting_loss(model: UDE, regularization_weight=0.0, pred_length=5) -> P.Problem:
    p = model.base_problem()
    p.loss_kind = P.LOSS_MULTI_SHOOT
    p.pred_length = int(pred_length)
    p.preroll_eps = 0.0 if model.multi else 1e-6      # LFC.jl:268 vs :646
    p.reg_weight = regularization_weight * model.weights["regularization"]
    return p.finalize()


def shooting_loss(model: UDE, regularization_weight=0.0) -> P.Problem:
    p = model.base_problem()
    p.loss_kind = P.LOSS_SHOOT
    p.shoot_first_scored = 2                           # LFC.jl:220-222, :572-574: `for i in 2:length(times)` over times[2:end]
    if model.multi:
        p.shoot_from_theta = False                     # LFC.jl:563: u0 = data[:, start]
        p.shoot_weight = 1.0
        p.reg_weight = regularization_weight * model.weights["regularization"]
    else:
        p.shoot_from_theta = True
        p.shoot_weight = model.T * model.weights["process"] / float(model.N) ** 2      # ProcessMSE with dt = 1
        p.reg_weight = model.weights["regularization"]                               # LFC.jl:225: not scaled by train!'s weight
    return p.finalize()


def smooth_derivs_batch(U, t, d=12, device=0):
    """Host pre-step of derivative matching (LFC.jl:161-171, :510-522) for MANY rows that share one time grid: U is (rows, n)
    -- every row one state variable of one series.  `RegularizationSmooth(u, t, d; alg = :gcv_svd)` + the derivative of its
    cubic-spline interpolant, restated in smoothing.py and evaluated on the GPU (csrc/ude_smooth.cu): the grid operators are
    built once on the host, one warp per row does the GCV minimisation.  Returns (smoothed, d/dt)."""
    U = np.asarray(U, dtype=np.float64); t = np.asarray(t, dtype=np.float64)
    n = t.shape[0]
    if n < 3:                   # the reference's smoother has nothing to penalise here (it would throw for n <= d)
        return U.copy(), (np.gradient(U, t, axis=1) if n > 1 else np.zeros_like(U))
    fit, der, _ = smoothing.reg_smooth_batch(U, t, d=d, device=device)
    return fit, der


def smooth_derivs(u, t, d=12):
    """one series: u is (d_state, n)"""
    return smooth_derivs_batch(u, t, d)


def derivative_matching_loss(model: UDE, regularization_weight=0.0, d=12, remove_ends=0) -> P.Problem:
    p = model.base_problem()
    p.loss_kind = P.LOSS_DERIV_MATCH
    p.remove_ends = int(remove_ends)
    us, du = np.zeros_like(model.data), np.zeros_like(model.data)
    starts = model.starts if model.multi else [0]
    lengths = model.lengths if model.multi else [model.data.shape[1]]
    # series that share a time grid are smoothed in one batch on the GPU (the operators depend on the grid only)
    groups = {}
    for s, n in zip(starts, lengths):
        groups.setdefault(np.asarray(model.times[s:s + n]).tobytes(), []).append((int(s), int(n)))
    dd = model.data.shape[0]
    for members in groups.values():
        s0, n = members[0]
        rows = np.concatenate([model.data[:, s:s + n] for s, _ in members], axis=0)
        fit, der = smooth_derivs_batch(rows, model.times[s0:s0 + n], d=d)
        for q, (s, _) in enumerate(members):
            us[:, s:s + n], du[:, s:s + n] = fit[dd * q: dd * (q + 1)], der[dd * q: dd * (q + 1)]
    p.dm_u, p.dm_dudt = us, du
    p.reg_weight = regularization_weight * model.weights["regularization"]
    return p.finalize()


def default_options(loss_function):
    """src/Optimizers.jl:242-258."""
    return {"conditional likelihood": dict(step_size=0.025, maxiter=500)}.get(loss_function, dict(step_size=0.05, maxiter=500))


def build_problem(model: UDE, loss_function, regularization_weight=0.0, loss_options=None, t_skip=None) -> P.Problem:
    o = dict(loss_options or {})
    if loss_function == "conditional likelihood":
        return conditional_likelihood(model, regularization_weight, o.get("observation_error", 0.025), o.get("process_error", 0.025), t_skip)
    if loss_function in ("derivative matching", "gradient matching", "smooth gradient matching"):
        return derivative_matching_loss(model, regularization_weight, d=o.get("d", 12), remove_ends=o.get("remove_ends", 0))
    if loss_function == "shooting":
        return shooting_loss(model, regularization_weight)
    if loss_function == "multiple shooting":
        return multiple_shooting_loss(model, regularization_weight, o.get("pred_length", 5))
    if loss_function == "default":
        return default_loss(model, t_skip)
    raise ValueError("Select a valid loss function. The CUDA hot path implements 'conditional likelihood', 'derivative matching', "
                     "'shooting', 'multiple shooting', 'marginal likelihood', 'neural gradient matching', 'spline gradient matching' "
                     "(and 'default' = model.loss_function).")


def adam_host(engine: Engine, theta, step_size, maxiter, verbose=False, callback=None):
    """Optimisers.jl Adam on the host, one ude_loss_grad call per iteration (what `train!` does through Optimization.solve)."""
    th = np.array(theta, dtype=np.float64, copy=True)
    m = np.zeros_like(th); v = np.zeros_like(th)
    trace = []
    for it in range(1, int(maxiter) + 1):
        L, g = engine.loss_grad(th)
        trace.append(L.copy())
        if verbose:
            print(round(float(L.sum()), 3), end=" ", flush=True)     # src/Optimizers.jl:480-483
        if callback is not None and callback(th, L):
            break
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        th -= step_size * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
    return th, np.array(trace)


def _interval_problem(model: UDE, t_from, t_to, series=0) -> P.Problem:
    """A state-space Problem (not finalized) whose series are independent one-interval mini-series [t_from[j], t_to[j]] of
    `model` (two columns each; the covariate knots of `series` repeated for every mini-series)."""
    d, n = model.d, len(t_from)
    m2 = copy.copy(model)
    m2.times = np.stack([np.asarray(t_from, dtype=np.float64), np.asarray(t_to, dtype=np.float64)], axis=1).reshape(-1)
    m2.data = np.zeros((d, 2 * n), order="F")
    m2.multi = True
    m2.starts = np.arange(n, dtype=np.int64) * 2
    m2.lengths = np.full(n, 2, dtype=np.int64)
    if model.cov:
        nc, off, kt, kx = model.cov
        blk = slice(series * nc, series * nc + nc + 1) if model.multi else slice(0, nc + 1)
        o = off[blk]
        seg_t, seg_x = kt[o[0]:o[-1]], kx[o[0]:o[-1]]
        m2.cov = (nc, np.concatenate([[0], np.tile(np.diff(o), n).cumsum()]).astype(np.int64), np.tile(seg_t, n), np.tile(seg_x, n))
    pr = m2.base_problem()
    pr.loss_kind = P.LOSS_STATE_SPACE
    return pr


def _interval_pm(model: UDE, pr: P.Problem, series=0) -> np.ndarray:
    """process_model for a mini-series problem of `_interval_problem`: the model's own, except that a per-series parameter
    vector (MultiCustomDerivatives' p.r[i]) becomes one entry per MINI-series, all equal to the entry of `series`."""
    pm = np.asarray(model.process_model, dtype=np.float64)
    if not pr.has_series_param:
        return pm
    n_own = 1 if not model.multi else len(model.starts)
    off = pm.shape[0] - n_own                    # the per-series vector is the tail of process_model (problem.default_layout)
    return np.concatenate([pm[:off], np.full(pr.n_series, pm[off + (series if model.multi else 0)])])


class NeuralGradientMatching:
    """neural_gradient_matching_loss(UDE, sigma, tau, regularization_weight, init_weights_layer_1)
    (src/LossFunctionConstructors.jl:764-846): an interpolating network uhat(t) = W2 tanh(W1 t + b1) + b2 (one hidden unit
    per observation, pre-fitted to the data with two Adam runs, :788-793) and
        loss = sum((u - uhat)^2) / sigma^2 + sum_t |duhat/dt(t) - rhs(uhat(t), pm, t)|^2 / tau^2 + w_reg (|W1|^2 + |W2|^2).
    The RHS term with its gradients w.r.t. uhat, duhat/dt and the process model is one batched evaluation on the GPU: every
    observation time is a two-column mini-series [uhat, uhat + h duhat/dt] under explicit Euler (h = 1), whose state-space
    residual is h (duhat/dt - rhs).  The interpolating network (N hidden units of one input) is host-side NumPy."""
    H_STEP = 1.0

    def __init__(self, model: UDE, sigma=0.1, tau=0.5, regularization_weight=0.0, init_weights_layer_1=5.0, seed=0, prefit=True):
        if model.multi:
            raise NotImplementedError("neural gradient matching is defined for single-series models in the reference")
        self.model, self.sigma, self.tau = model, float(sigma), float(tau)
        t = np.asarray(model.times, dtype=np.float64)
        self.t, self.N, d = t, t.shape[0], model.d
        rng = np.random.default_rng(seed)
        lim = np.sqrt(6.0 / (self.N + d))
        self.ip = dict(W1=np.full(self.N, float(init_weights_layer_1)), b1=-t * float(init_weights_layer_1),
                       W2=rng.uniform(-lim, lim, size=(d, self.N)).astype(np.float32).astype(np.float64), b2=np.zeros(d))
        if prefit:      # Adam(0.05) x 250, then Adam(0.025) x 500 on sum((u - uhat)^2)
            x = self.pack(self.ip)
            for step, iters in ((0.05, 250), (0.025, 500)):
                m = np.zeros_like(x); v = np.zeros_like(x)
                for it in range(1, iters + 1):
                    ip = self.unpack(x)
                    U, dU, H = self.interp(ip)
                    g = self.pack(self.interp_vjp(ip, H, -2.0 * (np.asarray(model.data) - U), np.zeros_like(U)))
                    m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
                    x = x - step * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
            self.ip = self.unpack(x)
        pr = _interval_problem(model, t, t + self.H_STEP)
        pr.solver, pr.substeps = P.EULER, 1
        pr.A, pr.B = np.zeros((d, d)), np.eye(d)
        pr.obs_scale, pr.proc_inv_dt = 0.0, False
        pr.proc_scale = 1.0 / (self.tau ** 2 * self.H_STEP ** 2)
        pr.reg_weight = float(regularization_weight)
        self.problem = pr.finalize()
        self.n_interp = self.pack(self.ip).shape[0]

    def pack(self, ip):
        return np.concatenate([ip["W1"], ip["b1"], ip["W2"].reshape(-1, order="F"), ip["b2"]])

    def unpack(self, x):
        N, d = self.N, self.model.d
        return dict(W1=x[:N], b1=x[N:2 * N], W2=x[2 * N:2 * N + d * N].reshape(d, N, order="F"), b2=x[2 * N + d * N:2 * N + d * N + d])

    def interp(self, ip):
        """uhat (d x N), duhat/dt (d x N) (dNNdt, :797-806) and the hidden activations."""
        H = np.tanh(ip["W1"][:, None] * self.t[None, :] + ip["b1"][:, None])
        return ip["W2"] @ H + ip["b2"][:, None], ip["W2"] @ (ip["W1"][:, None] * (1.0 - H * H)), H

    def interp_vjp(self, ip, H, GU, GD):
        S = 1.0 - H * H
        D = ip["W1"][:, None] * S
        W2b = GU @ H.T + GD @ D.T
        Hb = ip["W2"].T @ GU
        Db = ip["W2"].T @ GD
        W1b = np.sum(Db * S, axis=1)
        Hb = Hb + Db * ip["W1"][:, None] * (-2.0 * H)
        Zb = Hb * S
        return dict(W1=W1b + Zb @ self.t, b1=np.sum(Zb, axis=1), W2=W2b, b2=np.sum(GU, axis=1))

    def initial_parameters(self):
        """ComponentArray((UDE = UDE.parameters, interp = interp_params)) (:812)."""
        return np.concatenate([self.model.parameters, self.pack(self.ip)])

    def loss_grad(self, eng: Engine, x):
        d, N, h = self.model.d, self.N, self.H_STEP
        n_theta = self.model.parameters.shape[0]
        n_u = d * N
        theta, ip = x[:n_theta], self.unpack(x[n_theta:])
        U, dU, H = self.interp(ip)
        res = np.asarray(self.model.data) - U
        L_obs = float(np.sum(res * res)) / self.sigma ** 2
        cols = np.empty((d, 2 * N), order="F")
        cols[:, 0::2] = U; cols[:, 1::2] = U + h * dU
        L, g = eng.loss_grad(np.concatenate([cols.reshape(-1, order="F"), theta[n_u:]]))
        gc = g[: d * 2 * N].reshape(d, 2 * N, order="F")
        GU = gc[:, 0::2] + gc[:, 1::2] - (2.0 / self.sigma ** 2) * res
        GD = h * gc[:, 1::2]
        g_theta = np.zeros(n_theta)
        g_theta[n_u:] = g[d * 2 * N:]
        return L_obs + float(L[0]), np.concatenate([g_theta, self.pack(self.interp_vjp(ip, H, GU, GD))])


def neural_gradient_matching_(model: UDE, regularization_weight=0.0, loss_options=None, optim_options=None, verbose=False, device=0):
    """train!(model; loss_function = "neural gradient matching") (src/Optimizers.jl:453-459, :505, :562-563)."""
    o = dict(loss_options or {})
    opts = dict(default_options("neural gradient matching")); opts.update(optim_options or {})
    ng = NeuralGradientMatching(model, o.get("sigma", 0.1), o.get("tau", 0.5), regularization_weight, o.get("init_weights_layer_1", 5.0),
                                seed=o.get("seed", 0))
    x = ng.initial_parameters()
    m = np.zeros_like(x); v = np.zeros_like(x)
    trace = []
    with Engine(ng.problem, device=device) as eng:
        for it in range(1, int(opts["maxiter"]) + 1):
            L, g = ng.loss_grad(eng, x)
            trace.append(L)
            if verbose:
                print(round(L, 3), end=" ", flush=True)
            m = 0.9 * m + 0.1 * g
            v = 0.999 * v + 0.001 * g * g
            x = x - opts["step_size"] * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
    n_theta = model.parameters.shape[0]
    theta = x[:n_theta].copy()
    theta[: model.d * ng.N] = ng.interp(ng.unpack(x[n_theta:]))[0].reshape(-1, order="F")
    model.parameters = theta
    model.extra["loss_trace"] = np.array(trace)
    return None


class SplineGradientMatching:
    """spline_gradient_matching_loss(UDE, sigma, tau, regularization_weight, T) (src/LossFunctionConstructors.jl:888-931), the
    DEFAULT loss of train!(::UDE) (src/Optimizers.jl:354).  Parameters: knot values alpha (n_knots x d) and the UDE's theta.
        loss = sum(((u - uhat(alpha)) / sigma)^2) + w_reg * (|W1|^2 + |W2|^2)
             + N / That * sum_k |(alpha[k+1] - alpha[k]) / dt - rhs(alpha[k], pm, ts[k])|^2 / (Dt * tau^2),   k = 1 .. That - 1
    The last sum is a batched RHS + VJP: with explicit Euler as the "solver" it is exactly the state-space process term
    |alpha[k+1] - (alpha[k] + dt * rhs)|^2 / dt^2 on the knot grid, so it runs on the same fused kernels (UDE_EULER).  The
    observation term is a sparse linear map of alpha and stays on the host.  Quirks kept: knots run from t1 - 2dt to
    tN + 2dt; the interpolation weight (t - ts1)/(ts2 - ts1) multiplies the LEFT knot (:876-880); a data time that coincides
    with a knot is interpolated between its two neighbours (strict < and > in :865-866)."""

    def __init__(self, model: UDE, sigma=0.05 ** 2, tau=0.025 ** 2, regularization_weight=0.0, T=100):
        if model.multi:
            raise NotImplementedError("spline gradient matching is defined for single-series models in the reference")
        self.model, self.sigma = model, float(sigma)
        t = np.asarray(model.times, dtype=np.float64)
        N, d = t.shape[0], model.d
        T = int(T)
        dt = (t[-1] - t[0]) / T
        self.ts = (t[0] - 2 * dt) + dt * np.arange(T + 5)
        self.i1 = np.array([np.nonzero(self.ts < u)[0][-1] for u in t])
        self.i2 = np.array([np.nonzero(self.ts > u)[0][0] for u in t])
        self.w1 = (t - self.ts[self.i1]) / (self.ts[self.i2] - self.ts[self.i1])
        self.w2 = 1.0 - self.w1
        self.n_knots, self.That = T + 5, T + 4
        Dt = dt / ((t[-1] - t[0]) / N)
        m2 = copy.copy(model)
        m2.times = self.ts[: self.That].copy()
        m2.data = np.zeros((d, self.That), order="F")
        pr = m2.base_problem()
        pr.loss_kind, pr.solver, pr.substeps = P.LOSS_STATE_SPACE, P.EULER, 1
        pr.A, pr.B = np.zeros((d, d)), np.eye(d)
        pr.obs_scale, pr.proc_inv_dt = 0.0, False
        pr.proc_scale = (N / self.That) / (Dt * float(tau) ** 2) / dt ** 2
        pr.reg_weight = float(regularization_weight)
        self.problem = pr.finalize()
        self.n_alpha = self.n_knots * d

    def initial_parameters(self):
        """ComponentArray((alpha = zeros, UDE = UDE.parameters)) (:903-904): alpha is n_knots x d, column-major like Julia."""
        return np.concatenate([np.zeros(self.n_alpha), self.model.parameters])

    def states(self, alpha):
        """interp_states_spline (:876-882): d x N."""
        A = np.asarray(alpha).reshape(self.n_knots, self.model.d, order="F")
        return (A[self.i1] * self.w1[:, None] + A[self.i2] * self.w2[:, None]).T

    def loss_grad(self, eng: Engine, x):
        d, n_u = self.model.d, self.model.d * self.model.data.shape[1]
        alpha = x[: self.n_alpha].reshape(self.n_knots, d, order="F")
        theta = x[self.n_alpha:]
        # observation term on the host: O(N d)
        res = (np.asarray(self.model.data) - self.states(x[: self.n_alpha])) / self.sigma          # d x N
        L_obs = float(np.sum(res * res))
        g_alpha = np.zeros_like(alpha)
        gu = (-2.0 / self.sigma) * res.T                                                            # N x d: d L_obs / d uhat
        np.add.at(g_alpha, self.i1, gu * self.w1[:, None])
        np.add.at(g_alpha, self.i2, gu * self.w2[:, None])
        # dynamics term + regulariser on the GPU: theta_gpu = [alpha[1:That]' (d x That) | process_model]
        th = np.concatenate([alpha[: self.That].T.reshape(-1, order="F"), theta[n_u:]])
        L, g = eng.loss_grad(th)
        g_alpha[: self.That] += g[: d * self.That].reshape(d, self.That, order="F").T
        g_theta = np.zeros_like(theta)
        g_theta[n_u:] = g[d * self.That:]
        return L_obs + float(L[0]), np.concatenate([g_alpha.reshape(-1, order="F"), g_theta])


def spline_gradient_matching_(model: UDE, regularization_weight=0.0, loss_options=None, optim_options=None, verbose=False, device=0):
    """train!(model)'s default for single-series models (src/Optimizers.jl:354, :461-466, :565-566): Adam over (alpha, theta),
    then uhat <- interpolated states."""
    o = dict(loss_options or {})
    opts = dict(default_options("spline gradient matching")); opts.update(optim_options or {})
    sg = SplineGradientMatching(model, o.get("sigma", 0.05 ** 2), o.get("tau", 0.025 ** 2), regularization_weight, o.get("T", 100))
    x = sg.initial_parameters()
    m = np.zeros_like(x); v = np.zeros_like(x)
    trace = []
    with Engine(sg.problem, device=device) as eng:
        for it in range(1, int(opts["maxiter"]) + 1):
            L, g = sg.loss_grad(eng, x)
            trace.append(L)
            if verbose:
                print(round(L, 3), end=" ", flush=True)
            m = 0.9 * m + 0.1 * g
            v = 0.999 * v + 0.001 * g * g
            x = x - opts["step_size"] * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
    theta = x[sg.n_alpha:].copy()
    theta[: model.d * model.data.shape[1]] = sg.states(x[: sg.n_alpha]).reshape(-1, order="F")
    model.parameters = theta
    model.extra["loss_trace"] = np.array(trace)
    model.extra["spline_alpha"] = x[: sg.n_alpha].reshape(sg.n_knots, model.d, order="F").copy()
    return None


def marginal_likelihood_(model: UDE, regularization_weight=0.0, loss_options=None, optim_options=None, verbose=False, device=0):
    """train!(model; loss_function = "marginal likelihood") (src/Optimizers.jl:391-423, :500-560; MultiUDE :604-640): Adam over
    (process_model, P_nu factor) on the unscented-Kalman-filter likelihood, then uhat <- filtered states.  Returns
    (P_nu, Px) like the reference's `out`.  Options: process_error / observation_error (default 0.1), alpha, beta, kappa
    (defaults 1e-3, 2, 0)."""
    o = dict(loss_options or {})
    d = model.d
    Pnu0 = errors_to_matrix(o.get("process_error", 0.1), d)
    Peta = errors_to_matrix(o.get("observation_error", 0.1), d)
    alpha, beta, kappa = o.get("alpha", 1e-3), o.get("beta", 2.0), o.get("kappa", 0.0)
    opts = dict(default_options("marginal likelihood")); opts.update(optim_options or {})
    prob = default_loss(model)
    # the regulariser enters once (UDE, LFC.jl:119) or once per series plus once (MultiUDE, LFC.jl:441 and :457)
    reg = regularization_weight * model.weights["regularization"] * (_n_series(model) + 1 if model.multi else 1)
    F = np.linalg.cholesky(Pnu0)                       # Matrix(cholesky(P_nu).L)
    n_u, n_pm = d * model.data.shape[1], prob.n_pm
    x = np.concatenate([model.parameters[n_u:n_u + n_pm], F.reshape(-1, order="F")])
    m = np.zeros_like(x); v = np.zeros_like(x)
    theta = np.array(model.parameters, dtype=np.float64, copy=True)
    trace = []
    with Engine(prob, device=device) as eng:
        for it in range(1, int(opts["maxiter"]) + 1):
            theta[n_u:n_u + n_pm] = x[:n_pm]
            L, g, gF = eng.ukf_loss_grad(theta, x[n_pm:].reshape(d, d, order="F"), Peta, alpha, beta, kappa, reg)
            trace.append(L)
            if verbose:
                print(round(L, 3), end=" ", flush=True)
            gx = np.concatenate([g[n_u:n_u + n_pm], gF.reshape(-1, order="F")])
            m = 0.9 * m + 0.1 * gx
            v = 0.999 * v + 0.001 * gx * gx
            x = x - opts["step_size"] * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
        theta[n_u:n_u + n_pm] = x[:n_pm]
        F = x[n_pm:].reshape(d, d, order="F")
        xs, Ps = eng.ukf_filter(theta, F, Peta, alpha, beta, kappa)
    theta[:n_u] = xs.reshape(-1, order="F")
    model.parameters = theta
    model.extra["loss_trace"] = np.array(trace)
    return F @ F.T, Ps


def marginal_likelihood_many_(models, regularization_weight=0.0, loss_options=None, optim_options=None, verbose=False, device=0):
    """train!(...; loss_function = "marginal likelihood") for a LIST of models of one architecture (leave-future-out folds,
    ensemble members) in one many-model handle: every Adam iteration filters all of them in the same launches
    (ude_ukf_loss_grad on a many-model handle, one process model and one P_nu factor per model).  Adam is element-wise, so
    this equals len(models) separate marginal_likelihood_ runs.  Returns [(P_nu, Px), ...]; loss traces in model.extra."""
    from .cross_validation import stack_models
    o = dict(loss_options or {})
    d = models[0].d
    Pnu0 = errors_to_matrix(o.get("process_error", 0.1), d)
    Peta = errors_to_matrix(o.get("observation_error", 0.1), d)
    alpha, beta, kappa = o.get("alpha", 1e-3), o.get("beta", 2.0), o.get("kappa", 0.0)
    opts = dict(default_options("marginal likelihood")); opts.update(optim_options or {})
    regs = {regularization_weight * m.weights["regularization"] * (_n_series(m) + 1 if m.multi else 1) for m in models}
    if len(regs) != 1:
        raise ValueError("models filtered together share one regularisation weight")
    reg = regs.pop()
    big, theta = stack_models(models, "default")
    M, n_pm = big.n_models, big.n_pm
    col0 = np.asarray(big.model_col0, dtype=np.int64)
    base = np.asarray(big.theta_base, dtype=np.int64)
    pm_idx = (base + d * (col0[1:] - col0[:-1]))[:, None] + np.arange(n_pm)[None, :]           # (M, n_pm) into theta
    F0 = np.linalg.cholesky(Pnu0)
    x = np.concatenate([theta[pm_idx], np.tile(F0.reshape(-1, order="F"), (M, 1))], axis=1)    # (M, n_pm + d*d)
    m1 = np.zeros_like(x); v1 = np.zeros_like(x)
    trace = []
    as_factors = lambda xx: xx[:, n_pm:].reshape(M, d, d).transpose(0, 2, 1)                   # noqa: E731  (column-major rows)
    with Engine(big, device=device) as eng:
        for it in range(1, int(opts["maxiter"]) + 1):
            theta[pm_idx] = x[:, :n_pm]
            L, g, gF = eng.ukf_loss_grad(theta, as_factors(x), Peta, alpha, beta, kappa, reg)
            L, gF = np.atleast_1d(L), np.asarray(gF).reshape(M, d, d)
            trace.append(L.copy())
            if verbose:
                print(round(float(L.sum()), 3), end=" ", flush=True)
            gx = np.concatenate([g[pm_idx], gF.transpose(0, 2, 1).reshape(M, d * d)], axis=1)
            m1 = 0.9 * m1 + 0.1 * gx
            v1 = 0.999 * v1 + 0.001 * gx * gx
            x = x - opts["step_size"] * (m1 / (1 - 0.9 ** it)) / (np.sqrt(v1 / (1 - 0.999 ** it)) + 1e-8)
        theta[pm_idx] = x[:, :n_pm]
        Fs = as_factors(x)
        xs, Ps = eng.ukf_filter(theta, Fs, Peta, alpha, beta, kappa)
    trace = np.array(trace).reshape(-1, M)
    out = []
    for mi, mdl in enumerate(models):
        c0, c1 = int(col0[mi]), int(col0[mi + 1])
        th = theta[int(base[mi]): int(base[mi]) + d * (c1 - c0) + n_pm].copy()
        th[:d * (c1 - c0)] = xs[:, c0:c1].reshape(-1, order="F")
        mdl.parameters = th
        mdl.extra["loss_trace"] = trace[:, mi].copy()
        out.append((Fs[mi] @ Fs[mi].T, Ps[:, :, c0:c1]))
    return out


def kalman_filter_(model: UDE, sigma2, Pnu=None, Peta=None, alpha=1e-3, beta=2.0, kappa=0.0, verbose=False, maxiter=500, step_size=0.05,
                   device=0):
    """kalman_filter!(UDE::MultiUDE, sigma2; ...) (src/UnscentedKalmanFilter.jl:181-235): Adam over (process_model, p) with
    P_nu = diag(p.^2) (:117), the model's own regulariser once per series (:122), then uhat <- filtered states.  Quirk kept:
    the final smoothing pass and the returned matrix use diag(p), not diag(p.^2) (:156, :231).  Returns (P_nu, Px)."""
    d = model.d
    pv = np.full(d, float(sigma2)) if Pnu is None else np.asarray(Pnu, dtype=np.float64).copy()
    Pe = float(sigma2) * np.eye(d) if Peta is None else errors_to_matrix(Peta, d)
    prob = default_loss(model)
    reg = float(np.asarray(prob.reg_weight).reshape(-1)[0]) * _n_series(model)
    n_u, n_pm = d * model.data.shape[1], prob.n_pm
    x = np.concatenate([model.parameters[n_u:n_u + n_pm], pv])
    m = np.zeros_like(x); v = np.zeros_like(x)
    theta = np.array(model.parameters, dtype=np.float64, copy=True)
    trace = []
    with Engine(prob, device=device) as eng:
        for it in range(1, int(maxiter) + 1):
            theta[n_u:n_u + n_pm] = x[:n_pm]
            L, g, gF = eng.ukf_loss_grad(theta, np.diag(x[n_pm:]), Pe, alpha, beta, kappa, reg)
            trace.append(L)
            if verbose:
                print(round(L, 3), end=" ", flush=True)
            gx = np.concatenate([g[n_u:n_u + n_pm], np.diag(gF)])
            m = 0.9 * m + 0.1 * gx
            v = 0.999 * v + 0.001 * gx * gx
            x = x - step_size * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
        theta[n_u:n_u + n_pm] = x[:n_pm]
        Pn = np.diag(x[n_pm:])
        xs, Ps = eng.ukf_filter(theta, np.diag(np.sqrt(np.abs(x[n_pm:]))), Pe, alpha, beta, kappa)      # P_nu = diag(p) in the smoother
    theta[:n_u] = xs.reshape(-1, order="F")
    model.parameters = theta
    model.extra["loss_trace"] = np.array(trace)
    return Pn, Ps


def save_model_parameters(model: UDE, file):
    """src/LoadSaveModels.jl:1-5: the process-model parameters as a JSON array (flat ComponentVector order)."""
    import json
    with open(file, "w") as f:
        json.dump([float(x) for x in model.process_model], f)


def load_model_parameters_(model: UDE, file):
    """load_model_parameters! (src/LoadSaveModels.jl:7-11)."""
    import json
    with open(file) as f:
        vals = np.asarray(json.load(f), dtype=np.float64)
    if vals.shape != model.process_model.shape:
        raise ValueError("parameter file does not match the model's process_model")
    model.process_model[:] = vals


def _run_on_ranks(fns):
    """run one callable per rank on its own host thread (the fused exchange makes every evaluation a collective)"""
    import threading
    out, errs = [None] * len(fns), []

    def work(r):
        try:
            out[r] = fns[r]()
        except Exception as ex:      # noqa: BLE001
            errs.append(ex)

    ts = [threading.Thread(target=work, args=(r,)) for r in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    return out


def adam_multi_gpu(prob: P.Problem, theta, step_size, maxiter, devices, want_states=False):
    """Device-resident Adam on several GPUs of one box from ONE process: the series of a single-model problem are split into
    contiguous blocks (sharding.shard_problem), one handle and one host thread per GPU, ude_peer_export / ude_peer_attach
    by pointer; every iteration ends in the fused finalize + peer-memory exchange, so all ranks apply the identical
    update to their copy of the shared parameters.  Returns (theta, loss_trace[, states d x n_cols])."""
    from .sharding import shard_problem
    world = len(devices)
    theta = np.asarray(theta, dtype=np.float64)
    shards = [shard_problem(prob, theta, r, world) for r in range(world)]
    engines = []
    try:
        for r in range(world):
            engines.append(Engine(shards[r].problem, device=devices[r]))
        blobs = [e.peer_export(world) for e in engines]
        for r, e in enumerate(engines):
            e.peer_attach(blobs, world, r)
        res = _run_on_ranks([lambda r=r: engines[r].adam(shards[r].theta, step_size, maxiter) for r in range(world)])
        new = theta.copy()
        pmg = prob.pm_base(0)
        for r, sh in enumerate(shards):
            glo, ghi, llo = sh.uhat_slices[0]
            th_r = res[r][0]
            new[glo:ghi] = th_r[llo:llo + (ghi - glo)]
            n_u = ghi - glo
            if r == 0:
                new[pmg:pmg + sh.pm_shared] = th_r[n_u:n_u + sh.pm_shared]
            if prob.has_series_param:
                s0, s1 = sh.series
                new[pmg + sh.pm_shared + s0: pmg + sh.pm_shared + s1] = th_r[n_u + sh.pm_shared:]
        out = [new, res[0][1]]
        if want_states:
            parts = _run_on_ranks([lambda r=r: engines[r].forward(res[r][0]) for r in range(world)])
            out.append(np.concatenate(parts, axis=1))
        return tuple(out)
    finally:
        for e in engines:
            e.close()


def train_(model: UDE, loss_function=None, optimizer="ADAM", regularization_weight=0.0, verbose=True,
           loss_options=None, optim_options=None, t_skip=None, device=0, on_device: Optional[bool] = None, devices=None):
    """train!(model; ...) (src/Optimizers.jl:352-570, :578-742) on the CUDA hot path.  Mutates model.parameters.
    devices=(0, 1, ...): multi-series models with the ADAM optimiser are trained data-parallel over those GPUs (adam_multi_gpu).
    on_device=True keeps theta, Adam moments and the gradient on the GPU for all iterations;
    the default does so whenever no per-iteration printing is requested."""
    if loss_function is None:      # the reference's defaults: src/Optimizers.jl:354 (UDE), :579 (MultiUDE)
        loss_function = "derivative matching" if model.multi else "spline gradient matching"
    if optimizer not in ("ADAM", "BFGS"):
        raise ValueError("Select a valid optimization algorithm. Choose from 'ADAM' or 'BFGS'.")      # src/Optimizers.jl:524
    if loss_function == "neural gradient matching":
        if optimizer != "ADAM":
            raise ValueError("neural gradient matching: only optimizer='ADAM' is implemented on the CUDA path")
        return neural_gradient_matching_(model, regularization_weight, loss_options, optim_options, verbose, device)
    if loss_function == "spline gradient matching":
        if optimizer != "ADAM":
            raise ValueError("spline gradient matching: only optimizer='ADAM' is implemented on the CUDA path")
        return spline_gradient_matching_(model, regularization_weight, loss_options, optim_options, verbose, device)
    if loss_function == "marginal likelihood":
        if optimizer != "ADAM":
            raise ValueError("marginal likelihood: only optimizer='ADAM' is implemented on the CUDA path")
        return marginal_likelihood_(model, regularization_weight, loss_options, optim_options, verbose, device)
    prob = build_problem(model, loss_function, regularization_weight, loss_options, t_skip)
    opts = dict(default_options(loss_function)); opts.update(optim_options or {})
    on_device = (not verbose) if on_device is None else on_device
    if devices is not None and len(devices) > 1:
        if optimizer != "ADAM" or not model.multi or prob.n_models != 1:
            raise ValueError("devices=...: data-parallel training needs a multi-series model and optimizer='ADAM'")
        traj = loss_function in ("shooting", "multiple shooting")
        res = adam_multi_gpu(prob, model.parameters, opts["step_size"], opts["maxiter"], list(devices), want_states=traj)
        model.parameters = res[0]
        if traj:
            model.uhat[:, :] = res[2]
        elif loss_function in ("derivative matching", "gradient matching", "smooth gradient matching"):
            model.uhat[:, :] = prob.dm_u
        model.extra["loss_trace"] = res[1]
        return None
    with Engine(prob, device=device) as eng:
        if optimizer == "BFGS":
            # src/Optimizers.jl:506-510: Optim.BFGS on the same loss; here SciPy's BFGS driven by ude_loss_grad
            from scipy.optimize import minimize
            trace = []

            def fg(x):
                L, g = eng.loss_grad(x)
                trace.append(L.copy())
                return float(L.sum()), g

            res = minimize(fg, np.array(model.parameters, dtype=np.float64), jac=True, method="BFGS",
                           options=dict(maxiter=int((optim_options or {}).get("maxiter", 200))))
            theta, trace = res.x, np.array(trace)
        elif on_device:
            theta, trace = eng.adam(model.parameters, opts["step_size"], opts["maxiter"])
        else:
            theta, trace = adam_host(eng, model.parameters, opts["step_size"], opts["maxiter"], verbose=verbose)
        model.parameters = theta
        # states after training (src/Optimizers.jl:536-550): trajectories re-solved with the fitted parameters
        if loss_function in ("shooting", "multiple shooting"):
            model.uhat[:, :] = eng.forward(theta)
        elif loss_function in ("derivative matching", "gradient matching", "smooth gradient matching"):
            model.uhat[:, :] = prob.dm_u
    model.extra["loss_trace"] = trace
    return None


def gradient_descent_(model: UDE, step_size=0.05, maxiter=500, verbose=False, device=0):
    """gradient_descent!(model) (src/Optimizers.jl:12-39): Adam on model.loss_function (the default MSE state-space loss)."""
    return train_(model, "default", verbose=verbose, optim_options=dict(step_size=step_size, maxiter=maxiter), device=device)


def predictions(model: UDE, test_data=None, device=0):
    """predictions(UDE) -> one-step-ahead predictions F(uhat_{t-1}) for every column (src/ModelTesting.jl:181-195).
    predictions(UDE, test_data) -> (inits, obs, preds): one step ahead between consecutive rows of `test_data`, started
    from the observed values (src/ModelTesting.jl:245-260).  All intervals are one batch of segments on the GPU."""
    if test_data is None:
        prob = default_loss(model)
        with Engine(prob, device=device) as eng:
            return eng.forward(model.parameters)
    if model.multi:
        raise NotImplementedError("predictions(model, test_data) is defined for single-series models in the reference")
    cols = [model.time_column_name] + list(model.varnames)
    missing = [c for c in cols if c not in test_data.columns]
    if missing:
        raise ValueError(f"test data is missing columns {missing}")
    td = test_data.sort_values(model.time_column_name)
    m2 = copy.copy(model)
    m2.times = td[model.time_column_name].to_numpy(dtype=np.float64)
    m2.data = np.asfortranarray(td[list(model.varnames)].to_numpy(dtype=np.float64).T)
    pr = m2.base_problem()
    pr.loss_kind = P.LOSS_STATE_SPACE
    pr.finalize()
    theta = np.concatenate([np.asarray(m2.data).reshape(-1, order="F"), model.process_model])
    with Engine(pr, device=device) as eng:
        pred = eng.forward(theta)
    return np.asarray(m2.data)[:, :-1].copy(), np.asarray(m2.data)[:, 1:].copy(), pred[:, 1:].copy()


def forecast(model: UDE, u0, t0, times, series=0, device=0):
    """forecast(UDE, u0, t0, times) (src/ModelTesting.jl:541-571) with the extrapolation damping of init_forecast
    (src/ProcessModels.jl:2-25; l, extrap_rho from the constructor defaults): ude_forecast_chain, device-resident."""
    times = np.asarray(times, dtype=np.float64)
    assert np.all(times > t0), "t0 is greater than the first time point in times"
    d, n = model.d, times.shape[0]
    l, rho = model.extra.get("l", 0.25), model.extra.get("extrap_rho", 0.1)
    uh = model.uhat
    if model.multi:
        s0, ln = int(model.starts[series]), int(model.lengths[series])
        uh = uh[:, s0:s0 + ln]
    umax, umin, umean = uh.max(axis=1), uh.min(axis=1), uh.mean(axis=1)
    # one two-column series per forecast interval; column 0 of series j is filled with the running state
    tt = np.concatenate([[t0], times])
    pr = _interval_problem(model, tt[:-1], tt[1:], series)
    pr.finalize()
    theta = np.concatenate([np.zeros(d * 2 * n), _interval_pm(model, pr, series)])
    out = np.zeros((n, d + 1))
    with Engine(pr, device=device) as eng:     # the chain (predict, damp, next start) never leaves the device
        out[:, 1:] = eng.forecast_chain(theta, np.asarray(u0, dtype=np.float64), umax, umin, umean, l, rho).T
    out[:, 0] = times
    return out
