"""The derivative-matching smoother: RegularizationSmooth(u, t, d; alg = :gcv_svd) + spline derivative,
as restated in universaldiffeq.jl_b200/smoothing.py and run by csrc/ude_smooth.cu.

CPU part: the host-built grid operators and a NumPy transcription of the kernel's arithmetic against the independent
80-digit oracle (oracle/smooth_oracle_np.py) at 1e-10 -- including the reference's default order d = 12, where a plain
double-precision SVD alone would only reach 1e-7.  GPU part (-m gpu): the kernel itself against the same oracle and against
the transcription on a large batch."""
import numpy as np
import pytest
from scipy.interpolate import CubicSpline

import ude_b200 as ude
from oracle import smooth_oracle_np as so

sm = ude.smoothing


def kernel_transcription(u, t, d):
    """what ude_reg_smooth_kernel does for one row, in NumPy"""
    Vt, s, M = sm.grid_operators(t, d)
    z = Vt @ u; z2 = z * z; s2 = s * s; n = len(t)

    def terms(x):
        a = np.exp(2 * x) * s2; om = a / (1 + a); f = 1 / (1 + a)
        R = np.sum(om * om * z2); T = np.sum(om)
        return R, T, np.sum(4 * f * om * om * z2) * T - 2 * R * np.sum(2 * f * om)

    xlo, xhi = np.log(sm.LAMBDA_MIN), np.log(sm.LAMBDA_MAX)
    h = (xhi - xlo) / 63
    g = [n * terms(xlo + h * q)[0] / terms(xlo + h * q)[1] ** 2 for q in range(64)]
    bi = int(np.argmin(g))
    a, b, x = xlo + h * max(bi - 1, 0), xlo + h * min(bi + 1, 63), xlo + h * bi
    if terms(a)[2] < 0 < terms(b)[2]:
        for _ in range(200):
            if b - a <= 4e-16 * max(abs(a), abs(b)):
                break
            m = 0.5 * (a + b)
            a, b = (m, b) if terms(m)[2] < 0 else (a, m)
        x = 0.5 * (a + b)
    aa = np.exp(2 * x) * s2
    xv = u - Vt.T @ (aa / (1 + aa) * z)
    return xv, M @ xv, np.exp(x)


CASES = [(30, 12, False, 0), (24, 12, True, 1), (30, 2, True, 2), (12, 12, False, 3), (40, 4, True, 4)]


def _row(n, irregular, seed):
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.uniform(0.05, 0.15, size=n)) if irregular else np.linspace(0.0, 3.0, n)
    return t, np.sin(2 * t + seed) + 0.5 * np.cos(5 * t) + 0.1 * rng.normal(size=n)


def test_grid_operators():
    rng = np.random.default_rng(0)
    t = np.cumsum(rng.uniform(0.05, 0.15, size=17))
    # divided differences annihilate polynomials of degree < d and differentiate t^d to d!
    for d in (1, 2, 5):
        D = sm.derivative_matrix(t, d)
        assert D.shape == (17 - d, 17)
        assert np.max(np.abs(D @ t ** (d - 1))) < 1e-6 * np.max(np.abs(D))
        assert np.allclose(D @ t ** d, float(np.prod(np.arange(1, d + 1))), rtol=1e-8)
    # knot derivatives of the natural cubic spline (DataInterpolations.CubicSpline: z_1 = z_n = 0)
    M = sm.spline_derivative_matrix(t)
    x = rng.normal(size=17)
    assert np.max(np.abs(M @ x - CubicSpline(t, x, bc_type="natural")(t, 1))) < 1e-11
    Vt, s, _ = sm.grid_operators(t, 12)
    assert np.max(np.abs(Vt @ Vt.T - np.eye(Vt.shape[0]))) < 1e-13
    for k in range(12):                       # the penalised directions are orthogonal to every polynomial of degree < 12
        p = (2 * (t - t[0]) / (t[-1] - t[0]) - 1) ** k
        assert np.max(np.abs(Vt @ p)) < 1e-12 * np.linalg.norm(p)


@pytest.mark.parametrize("n,d,irregular,seed", CASES)
def test_kernel_arithmetic_matches_oracle(n, d, irregular, seed):
    t, u = _row(n, irregular, seed)
    x1, d1, l1 = kernel_transcription(u, t, d)
    x2, d2, l2 = so.smooth(u, t, d)
    assert np.max(np.abs(x1 - x2)) < 1e-10 * np.max(np.abs(x2))
    assert np.max(np.abs(d1 - d2)) < 1e-10 * np.max(np.abs(d2))
    if d <= 4:                                # for d = 12 GCV is flat in lambda (every penalised direction is fully damped)
        assert abs(l1 - l2) < 1e-9 * l2


@pytest.mark.gpu
def test_gpu_smoother_matches_oracle_and_transcription():
    for n, d, irregular, seed in CASES:
        t, u = _row(n, irregular, seed)
        rng = np.random.default_rng(seed + 100)
        U = np.stack([u] + [np.sin(rng.uniform(1, 3) * t + q) + 0.1 * rng.normal(size=n) for q in range(300)])
        X, DX, lam = sm.reg_smooth_batch(U, t, d=d)
        x2, d2, _ = so.smooth(u, t, d)
        assert np.max(np.abs(X[0] - x2)) < 1e-10 * np.max(np.abs(x2))
        assert np.max(np.abs(DX[0] - d2)) < 1e-10 * np.max(np.abs(d2))
        for q in (1, 57, 300):
            xr, dr, lr = kernel_transcription(U[q], t, d)
            assert np.max(np.abs(X[q] - xr)) < 1e-11 * np.max(np.abs(xr)) and np.max(np.abs(DX[q] - dr)) < 1e-10 * np.max(np.abs(dr))
            if d <= 4:
                assert abs(lam[q] - lr) < 1e-9 * lr


@pytest.mark.gpu
def test_derivative_matching_uses_the_device_smoother():
    df = ude.LotkaVolterra(plot=False, datasize=40)
    m = ude.NODE(df, hidden_units=6, seed=2)
    prob = ude.training.derivative_matching_loss(m, d=12)
    X, DX, _ = sm.reg_smooth_batch(m.data, m.times, d=12)
    assert np.array_equal(prob.dm_u, X) and np.array_equal(prob.dm_dudt, DX)
    # smooth data is reproduced and differentiated accurately
    t = np.linspace(0, 2, 50)
    X, DX, _ = sm.reg_smooth_batch(np.stack([np.sin(2 * t), t ** 3]), t, d=4)
    assert np.max(np.abs(X[0] - np.sin(2 * t))) < 1e-3 and np.max(np.abs(DX[0, 3:-3] - 2 * np.cos(2 * t[3:-3]))) < 2e-2
    assert np.max(np.abs(X[1] - t ** 3)) < 1e-9       # a cubic is in the null space of the 4th-order penalty
