"""NumPy restatement of the reference's unscented-Kalman-filter marginal likelihood.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (no Julia in this image; the reference's tests hold no golden values for this loss).  Follows
  /root/reference/src/UnscentedKalmanFilter.jl:2-8     sigma_points  (upper Cholesky factor of (lambda+L) Px, weights)
  /root/reference/src/UnscentedKalmanFilter.jl:10-55   ukf_update    (2L+1 predicts, mean, covariance, gain, log-likelihood with
                                                                      the literal 3.14159 in the normalising constant)
  /root/reference/src/UnscentedKalmanFilter.jl:59-87   ukf_smoothing / ukf_likelihood (x1 = y[:,1], Px = P_eta, steps 2..T)
  /root/reference/src/LossFunctionConstructors.jl:103-127 (UDE)  :427-476 (MultiUDE: regulariser once per series plus once)
with the one-interval predict F(u, t, dt) of the fixed-step solver (oracle/ude_oracle_np.flow).  dtype-generic: with a
complex theta the loss is analytic in it (no conj, |.| or comparisons on values), so `complex_step_grad` yields the
gradient that the CUDA path's hand-written adjoint is tested against.  Pure Python loops: small cases only.
"""
from __future__ import annotations

import numpy as np

from . import ude_oracle_np as onp


def chol_upper(A):
    """U upper triangular with U^T U = A, reading only the upper triangle of A (Julia: cholesky(Hermitian(A)).U)."""
    n = A.shape[0]
    U = np.zeros((n, n), dtype=A.dtype)
    for i in range(n):
        s = A[i, i]
        for k in range(i):
            s = s - U[k, i] * U[k, i]
        U[i, i] = np.sqrt(s)
        for j in range(i + 1, n):
            s = A[i, j]
            for k in range(i):
                s = s - U[k, i] * U[k, j]
            U[i, j] = s / U[i, i]
    return U


def _inv_det(S):
    """inverse and determinant of a small dense matrix: LAPACK for double / complex double, Gauss-Jordan with partial pivoting
    for everything else (np.longdouble: the extended-precision runs that size the alpha = 1e-3 tolerance)"""
    if S.dtype in (np.float64, np.complex128):
        return np.linalg.inv(S), np.linalg.det(S)
    n = S.shape[0]
    A = np.concatenate([S.copy(), np.eye(n, dtype=S.dtype)], axis=1)
    det = S.dtype.type(1)
    for c in range(n):
        piv = c + int(np.argmax(np.abs(A[c:, c])))
        if piv != c:
            A[[c, piv]] = A[[piv, c]]; det = -det
        det = det * A[c, c]
        A[c] = A[c] / A[c, c]
        for r in range(n):
            if r != c:
                A[r] = A[r] - A[r, c] * A[c]
    return A[:, n:], det


def weights(L, alpha, beta, kappa):
    lam = alpha ** 2 * (L - kappa) - L
    return lam + L, lam / (lam + L), lam / (lam + L) + (1 - alpha ** 2 + beta), 1.0 / (2 * (L + lam))


def ukf_update(y, x, Px, f, Pnu, Peta, L, alpha, beta, kappa):
    c, Wm0, Wc0, Wi = weights(L, alpha, beta, kappa)
    U = chol_upper(c * Px)
    pts = [f(x)] + [f(x + U[i, :]) for i in range(L)] + [f(x - U[i, :]) for i in range(L)]
    m = Wm0 * pts[0]
    for q in pts[1:]:
        m = m + Wi * q
    P1 = Wc0 * np.outer(pts[0] - m, pts[0] - m)
    for q in pts[1:]:
        P1 = P1 + Wi * np.outer(q - m, q - m)
    P1 = P1 + Pnu
    S = P1 + Peta
    Sinv, detS = _inv_det(S)
    K = P1 @ Sinv
    r = y - m
    xn = m + K @ r
    Pn = P1 - K @ P1
    ll = -0.5 * (r @ (Sinv @ r)) - 0.5 * np.log(detS) - L / 2 * np.log(2 * 3.14159)
    return xn, Pn, ll


def filter_series(p, pm, s, Pnu, Peta, alpha, beta, kappa, dtype=float, causal=False):
    """(sum of log-likelihood terms, filtered states d x n, covariances d x d x n) of series s.

    Prediction window of filter step t (1-based t = 2..T in the reference).  ukf_likelihood / ukf_smoothing call
    `ukf_update(y[:,t], x, Px, times[t], dt, ...)` with dt = times[t] - times[t-1] (src/UnscentedKalmanFilter.jl:67, :83) and
    ukf_update hands (t, dt) to predict (:93, multi :121): the sigma points are integrated over [times[t], times[t] + dt],
    i.e. the window STARTS at the END of the observation interval.  For an autonomous right-hand side this equals the causal
    window [times[t-1], times[t]]; with covariates or a time input it does not.  causal=False (default) reproduces the
    reference; causal=True is the window one would expect, kept as an opt-in (UDE_OPT_UKF_CAUSAL_INTERVAL)."""
    d = p.d
    c0, n = int(p.starts[s]), int(p.lengths[s])
    y = p.data[:, c0:c0 + n]
    tt = p.times[c0:c0 + n]
    x = y[:, 0].astype(dtype)
    Px = np.asarray(Peta, dtype=dtype)
    xs = np.zeros((d, n), dtype=dtype); Ps = np.zeros((d, d, n), dtype=dtype)
    xs[:, 0] = x
    ll = 0
    for t in range(1, n):
        ta, tb = (tt[t - 1], tt[t]) if causal else (tt[t], tt[t] + (tt[t] - tt[t - 1]))
        f = lambda u, t0=ta, t1=tb: onp.flow(p, pm, s, s, u, t0, t1)      # noqa: E731
        x, Px, l1 = ukf_update(y[:, t], x, Px, f, Pnu, Peta, d, alpha, beta, kappa)
        ll = ll + l1
        xs[:, t] = x; Ps[:, :, t] = Px
    return ll, xs, Ps


def _model_range(p, model):
    """(offset of the process model inside theta, first series, end series) of `model` (None: single-model problem)"""
    if model is None:
        return p.d * p.n_cols, 0, p.n_series
    return (int(p.theta_base[model]) + p.d * int(p.model_col0[model + 1] - p.model_col0[model]),
            int(p.model_series0[model]), int(p.model_series0[model + 1]))


def loss(p, theta, pnu, Peta, alpha=1e-3, beta=2.0, kappa=0.0, reg_weight=0.0, reg_mult=1.0, causal=False, model=None):
    """-sum_series loglik + reg_mult * reg_weight * (|W1|^2 + |W2|^2).  theta = [uhat | process_model] of a single-model
    problem (uhat does not enter); pnu = d x d factor with P_nu = pnu pnu^T.  reg_mult = 1 (UDE) or n_series + 1 (MultiUDE).
    model = m: the loss of model m of a many-model problem (its own series, its own theta block; pnu is that model's factor)."""
    theta = np.asarray(theta); pnu = np.asarray(pnu)
    dtype = complex if (np.iscomplexobj(theta) or np.iscomplexobj(pnu)) else (np.longdouble if theta.dtype == np.longdouble else float)
    if dtype is np.longdouble:
        Peta = np.asarray(Peta, dtype=np.longdouble); pnu = pnu.astype(np.longdouble)
    d = p.d
    off, s0, s1 = _model_range(p, model)
    pm = theta[off: off + p.n_pm]
    Lf = pnu.reshape(d, d, order="F")
    Pnu = Lf @ Lf.T
    tot = 0
    for s in range(s0, s1):
        tot = tot - filter_series(p, pm, s, Pnu, Peta, alpha, beta, kappa, dtype, causal=causal)[0]
    w1 = pm[p.off_w1:p.off_w1 + p.nn_hidden * p.nn_in]; w2 = pm[p.off_w2:p.off_w2 + p.nn_out * p.nn_hidden]
    return tot + reg_mult * reg_weight * (np.sum(w1 * w1) + np.sum(w2 * w2))


def complex_step_grad(p, theta, pnu, Peta, eps=1e-30, **kw):
    """(d loss / d process_model, d loss / d pnu) by complex step."""
    d = p.d
    n_u = _model_range(p, kw.get("model"))[0]
    g_pm = np.zeros(p.n_pm); g_nu = np.zeros(d * d)
    th = np.asarray(theta, dtype=complex); pv = np.asarray(pnu, dtype=complex).reshape(-1, order="F")
    for k in range(p.n_pm):
        t2 = th.copy(); t2[n_u + k] += 1j * eps
        g_pm[k] = np.imag(loss(p, t2, pv.reshape(d, d, order="F"), Peta, **kw)) / eps
    for k in range(d * d):
        p2 = pv.copy(); p2[k] += 1j * eps
        g_nu[k] = np.imag(loss(p, th, p2.reshape(d, d, order="F"), Peta, **kw)) / eps
    return g_pm, g_nu
