"""UKF marginal likelihood: the NumPy restatement against a known answer (linear dynamics: the unscented
transform is exact, so the loss must equal the textbook Kalman filter's), and the reverse sweep that
csrc/ude_ukf.cu implements -- restated here in NumPy with complex-step Jacobians standing in for the segment kernel's
VJP -- against the complex-step gradient of the whole loss."""
import numpy as np
import pytest

import cases
from oracle import ude_oracle_np as onp
from oracle import ukf_oracle_np as ukf


def _case(rhs="node_d2", hidden=4, lengths=(6,), solver=0, seed=0):
    prob, theta = cases.make_case(rhs, "cond_lik", solver, hidden=hidden, lengths=lengths, seed=seed)
    d = prob.d
    rng = np.random.default_rng(seed + 1)
    Lf = 0.3 * np.eye(d) + 0.05 * rng.normal(size=(d, d))
    Peta = 0.1 * np.eye(d) + 0.02 * np.ones((d, d))
    return prob, theta, Lf, Peta


def test_linear_dynamics_equal_kalman_filter():
    """zero network + linear decay term: f(u) = exp(-rho dt) u per interval (up to the solver's order) -- use the exact
    one-step map of the fixed-step solver as the KF transition."""
    prob, theta = cases.make_case("nn_decay_d1", "cond_lik", 0, hidden=3, lengths=(9,))
    d = prob.d
    n_u = d * prob.n_cols
    pm = theta[n_u:].copy()
    pm[prob.off_w2:prob.off_w2 + prob.nn_out * prob.nn_hidden] = 0.0      # W2 = 0
    pm[prob.off_b2:prob.off_b2 + prob.nn_out] = 0.0
    theta[n_u:] = pm
    Lf = np.array([[0.2]]); Peta = np.array([[0.05]])
    for alpha in (1e-3, 1.0):
        got = ukf.loss(prob, theta, Lf, Peta, alpha=alpha)
        # textbook scalar KF
        y, tt = prob.data[0], prob.times
        x, P, ll = y[0], Peta[0, 0], 0.0
        for t in range(1, len(y)):
            a = onp.flow(prob, pm, 0, 0, np.array([1.0]), tt[t - 1], tt[t])[0]       # linear map: f(u) = a u
            m, P1 = a * x, a * a * P + 0.04
            S = P1 + Peta[0, 0]
            K = P1 / S
            ll += -0.5 * (y[t] - m) ** 2 / S - 0.5 * np.log(S) - 0.5 * np.log(2 * 3.14159)
            x, P = m + K * (y[t] - m), P1 - K * P1
        tol = 1e-12 if alpha == 1.0 else 1e-5          # alpha = 1e-3: weights of +-1e6 cancel ~6 digits (reference default)
        assert abs(got + ll) < tol * max(1.0, abs(ll)), (alpha, got, -ll)


def chol_upper_adjoint(U, Ubar):
    n = U.shape[0]
    Ub = Ubar.copy(); Ab = np.zeros_like(U)
    for i in range(n - 1, -1, -1):
        for j in range(n - 1, i, -1):
            sb = Ub[i, j] / U[i, i]
            Ub[i, i] -= Ub[i, j] * U[i, j] / U[i, i]
            Ab[i, j] = sb
            for k in range(i):
                Ub[k, i] -= sb * U[k, j]
                Ub[k, j] -= sb * U[k, i]
        sb = Ub[i, i] / (2 * U[i, i])
        Ab[i, i] = sb
        for k in range(i):
            Ub[k, i] -= 2 * sb * U[k, i]
    return Ab


def reverse_sweep(prob, theta, Lf, Peta, alpha, beta, kappa):
    """loss and gradient (process_model, Lf) by the hand-derived adjoint; f's Jacobians by complex step."""
    d = prob.d
    n_u = d * prob.n_cols
    pm = theta[n_u:n_u + prob.n_pm]
    Pnu = Lf @ Lf.T
    c, Wm0, Wc0, Wi = ukf.weights(d, alpha, beta, kappa)
    w_m = np.array([Wm0] + [Wi] * (2 * d)); w_c = np.array([Wc0] + [Wi] * (2 * d))
    g_pm = np.zeros(prob.n_pm); Pnu_bar = np.zeros((d, d)); total = 0.0
    eps = 1e-30
    for s in range(prob.n_series):
        c0, n = int(prob.starts[s]), int(prob.lengths[s])
        y, tt = prob.data[:, c0:c0 + n], prob.times[c0:c0 + n]
        x, P = y[:, 0].copy(), Peta.copy()
        tape = []
        for t in range(1, n):
            U = ukf.chol_upper(c * P)
            chi = [x] + [x + U[i] for i in range(d)] + [x - U[i] for i in range(d)]
            Y = [onp.flow(prob, pm, s, s, q, tt[t - 1], tt[t]) for q in chi]
            m = sum(w * q for w, q in zip(w_m, Y))
            E = [q - m for q in Y]
            P1 = sum(w * np.outer(e, e) for w, e in zip(w_c, E)) + Pnu
            S = P1 + Peta; G = np.linalg.inv(S); K = P1 @ G; r = y[:, t] - m
            total -= -0.5 * r @ G @ r - 0.5 * np.log(np.linalg.det(S)) - d / 2 * np.log(2 * 3.14159)
            tape.append((t, U, chi, E, P1, G, K, r))
            x, P = m + K @ r, P1 - K @ P1
        xb, Pb = np.zeros(d), np.zeros((d, d))
        for t, U, chi, E, P1, G, K, r in reversed(tape):
            llb = -1.0 * -1.0          # loss = -ll
            llb = 1.0
            Kb = np.outer(xb, r) - Pb @ P1.T
            rb = K.T @ xb + llb * 0.5 * (G + G.T) @ r
            mb = xb - rb
            P1b = Pb - K.T @ Pb + Kb @ G.T
            Gb = P1.T @ Kb + llb * 0.5 * np.outer(r, r)
            Sb = -G.T @ Gb @ G.T + llb * 0.5 * G.T
            P1b = P1b + Sb
            Pnu_bar += P1b
            Eb = [w * (P1b + P1b.T) @ e for w, e in zip(w_c, E)]
            mb = mb - sum(Eb)
            Yb = [eb + w * mb for eb, w in zip(Eb, w_m)]
            chib = []
            for q, lam in zip(chi, Yb):      # VJP of the one-interval flow (the segment kernel's job on the GPU)
                Ju = np.array([np.imag(onp.flow(prob, pm.astype(complex), s, s, q + 1j * eps * np.eye(d)[a], tt[t - 1], tt[t])) / eps
                               for a in range(d)])
                chib.append(Ju @ lam)
                for k in range(prob.n_pm):
                    p2 = pm.astype(complex); p2[k] += 1j * eps
                    g_pm[k] += (np.imag(onp.flow(prob, p2, s, s, q.astype(complex), tt[t - 1], tt[t])) / eps) @ lam
            xb = sum(chib)
            Ub = np.array([chib[1 + i] - chib[1 + d + i] for i in range(d)])
            Pb = c * chol_upper_adjoint(U, Ub)
    return total, g_pm, (Pnu_bar + Pnu_bar.T) @ Lf


@pytest.mark.parametrize("alpha", [1.0, 0.05])
def test_reverse_sweep_matches_complex_step(alpha):
    prob, theta, Lf, Peta = _case("node_d2", hidden=3, lengths=(5, 3))
    L0 = ukf.loss(prob, theta, Lf, Peta, alpha=alpha)
    g_pm, g_nu = ukf.complex_step_grad(prob, theta, Lf, Peta, alpha=alpha)
    L1, h_pm, h_nu = reverse_sweep(prob, theta, Lf, Peta, alpha, 2.0, 0.0)
    assert abs(L1 - L0) < 1e-12 * abs(L0)
    assert cases.relerr(h_pm, g_pm) < 1e-9 / alpha ** 2
    assert cases.relerr(h_nu.reshape(-1, order="F"), g_nu) < 1e-9 / alpha ** 2


def _golden():
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden_ukf as mg
    return mg, np.load(os.path.join(here, "golden", "golden_ukf_twin.npz"))


def test_ukf_twin_matches_committed_fixtures():
    """tests/golden/golden_ukf_twin.npz (written by make_golden_ukf.py): loss and both gradients of six cases."""
    mg, gold = _golden()
    for rhs, hidden, lengths, solver, alpha in mg.CASES:
        prob, theta, F, Peta = mg.setup(rhs, hidden, lengths, solver)
        key = f"{rhs}|{hidden}|{'-'.join(map(str, lengths))}|{solver}|{alpha}"
        L = ukf.loss(prob, theta, F, Peta, alpha=alpha, reg_weight=1e-3, reg_mult=2.0)
        tol = 1e-12 if alpha == 1.0 else 1e-6
        assert abs(L - gold[key + "|loss"][0]) < tol * abs(L)
        if rhs == "node_d2" and len(lengths) == 1:          # one gradient re-derived here; the rest are checked on the GPU
            g_pm, g_F = ukf.complex_step_grad(prob, theta, F, Peta, alpha=alpha, reg_weight=1e-3, reg_mult=2.0)
            assert cases.relerr(g_pm, gold[key + "|g_pm"]) < 1e-12 and cases.relerr(g_F, gold[key + "|g_F"]) < 1e-12


def test_prediction_window_is_the_reference_one():
    """ukf_likelihood passes times[t] -- the END of the observation interval -- as the start time of predict
    (src/UnscentedKalmanFilter.jl:83, :93): the sigma points are integrated over [times[t], times[t] + dt].  The twin does
    the same by default; `causal=True` is the window [times[t-1], times[t]].  Equal for an autonomous right-hand side,
    different with a time input or covariates -- and the default must be the reference's window, written out here."""
    for rhs, autonomous in (("node_d2", True), ("node_time_d3", False), ("node_d4x1", False)):
        prob, theta, Lf, Peta = _case(rhs, hidden=4, lengths=(5, 3))
        a = ukf.loss(prob, theta, Lf, Peta, alpha=1.0)
        b = ukf.loss(prob, theta, Lf, Peta, alpha=1.0, causal=True)
        assert (abs(a - b) < 1e-13 * abs(a)) == autonomous, (rhs, a, b)
        # the default, step by step, with the window written as the reference computes it
        d = prob.d
        pm = theta[d * prob.n_cols: d * prob.n_cols + prob.n_pm]
        tot = 0.0
        for s in range(prob.n_series):
            c0, n = int(prob.starts[s]), int(prob.lengths[s])
            y, tt = prob.data[:, c0:c0 + n], prob.times[c0:c0 + n]
            x, P = y[:, 0].copy(), np.asarray(Peta, dtype=float)
            for t in range(1, n):
                dt = tt[t] - tt[t - 1]
                f = lambda u, t0=tt[t], t1=tt[t] + dt: onp.flow(prob, pm, s, s, u, t0, t1)      # noqa: E731
                x, P, ll = ukf.ukf_update(y[:, t], x, P, f, Lf @ Lf.T, Peta, d, 1.0, 2.0, 0.0)
                tot -= ll
        assert abs(tot - a) < 1e-13 * abs(a), rhs


def test_alpha_conditioning_extended_precision():
    """How well CAN two FP64 evaluations of the reference's UKF loss agree?  The twin in np.longdouble (64-bit mantissa) is the
    yardstick: at alpha = 1 the FP64 twin matches it to rounding; at the reference default alpha = 1e-3 the sigma-point weights
    (~ +-1e6) cancel and the FP64 value sits ~1e-11 away -- growing like 1 / alpha^2.  This is what sizes the tolerance of the
    alpha = 1e-3 GPU tests (1e-9 / 5e-8), and shows 1e-10 on the gradient is not attainable in FP64 for that alpha."""
    prob, theta, Lf, Peta = _case("node_d2", hidden=5, lengths=(8, 5))
    err = {}
    for alpha in (1.0, 1e-2, 1e-3):
        a = ukf.loss(prob, theta, Lf, Peta, alpha=alpha)
        b = float(ukf.loss(prob, theta.astype(np.longdouble), Lf, Peta, alpha=alpha))
        err[alpha] = abs(a - b) / abs(b)
    assert err[1.0] < 1e-15 and err[1e-2] < 1e-11 and err[1e-3] < 1e-9
    assert err[1e-3] > 30 * max(err[1.0], 1e-16)          # the cancellation is real: well above FP64 rounding
    # gradient: complex step in double vs in extended precision, three process-model entries
    n_u = prob.d * prob.n_cols
    eps = 1e-30
    for k in (0, 7, prob.n_pm - 1):
        t64 = theta.astype(complex); t64[n_u + k] += 1j * eps
        tld = theta.astype(np.clongdouble); tld[n_u + k] += np.clongdouble(1j * eps)
        g64 = np.imag(ukf.loss(prob, t64, Lf, Peta, alpha=1e-3)) / eps
        gld = float(np.imag(ukf.loss(prob, tld, Lf.astype(np.clongdouble), np.asarray(Peta, dtype=np.clongdouble), alpha=1e-3)) / eps)
        assert abs(g64 - gld) < 1e-7 * max(abs(gld), 1e-3)



def test_model_argument_selects_one_model_of_a_stacked_problem():
    """loss(..., model=m) on a many-model problem = the same series and process model as a single-model problem of their own
    (what tests/test_gpu_ukf.py::test_ukf_many_models checks the CUDA many-model filter against)."""
    many, th = cases.make_case("node_time_d3", "cond_lik", 0, hidden=4, lengths=(5, 3, 4, 6), n_models=2, seed=4)
    d = many.d
    rng = np.random.default_rng(2)
    Peta = 0.1 * np.eye(d)
    for m in range(2):
        Lf = 0.3 * np.eye(d) + 0.05 * rng.normal(size=(d, d))
        got = ukf.loss(many, th, Lf, Peta, alpha=1.0, reg_weight=1e-3, model=m)
        off, s0, s1 = ukf._model_range(many, m)
        pm = th[off:off + many.n_pm]
        tot = 0.0
        for s in range(s0, s1):
            tot -= ukf.filter_series(many, pm, s, Lf @ Lf.T, Peta, 1.0, 2.0, 0.0)[0]
        w1 = pm[many.off_w1:many.off_w1 + many.nn_hidden * many.nn_in]; w2 = pm[many.off_w2:many.off_w2 + many.nn_out * many.nn_hidden]
        assert abs(got - (tot + 1e-3 * (w1 @ w1 + w2 @ w2))) < 1e-13 * abs(got)
        g_pm, g_F = ukf.complex_step_grad(many, th, Lf, Peta, alpha=1.0, model=m)
        # finite-difference spot check of two entries of each gradient (the complex step perturbs the right theta block)
        for k in (0, many.n_pm - 1):
            t2 = th.copy(); t2[off + k] += 1e-6
            t3 = th.copy(); t3[off + k] -= 1e-6
            fd = (ukf.loss(many, t2, Lf, Peta, alpha=1.0, model=m) - ukf.loss(many, t3, Lf, Peta, alpha=1.0, model=m)) / 2e-6
            assert abs(fd - g_pm[k]) < 1e-5 * max(1.0, abs(g_pm[k]))
        other = 1 - m
        off_o = ukf._model_range(many, other)[0]
        t4 = th.copy(); t4[off_o:off_o + many.n_pm] += 0.1               # the other model's parameters do not enter
        assert ukf.loss(many, t4, Lf, Peta, alpha=1.0, reg_weight=1e-3, model=m) == got
