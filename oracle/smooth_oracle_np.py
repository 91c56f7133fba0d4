"""Dense NumPy / mpmath check of the derivative-matching smoother.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's smoother is third-party Julia (DataInterpolations.RegularizationSmooth over
RegularizationTools, /root/reference/Project.toml:47,65; called at LFC.jl:165, :513) and cannot run here.  This restates the
published algorithm without sharing code or formulation with universaldiffeq.jl_b200/smoothing.py / csrc/ude_smooth.cu:
  * D by the textbook divided-difference recursion, row by row, in 80-digit arithmetic (mpmath);
  * the symmetric eigen-decomposition of D'D (mp.eigsy) instead of an SVD of D: x(lambda) = Q diag(1 / (1 + lambda^2 E)) Q' u,
    GCV(lambda) = n |u - x|^2 / tr(I - H)^2;
  * lambda = argmin GCV over [1e-4, 1e3] (RegularizationTools.solve defaults, Brent on log10 lambda there): a 64-point scan and
    a golden-section search inside the best bracket, still in 80 digits -- a flat minimum is located to ~1e-8 in double
    precision, to ~1e-30 here, so the comparison with the CUDA kernel (which solves dGCV = 0) can use a tight bar;
  * derivative of the natural cubic spline by scipy.interpolate.CubicSpline(bc_type='natural')."""
import mpmath as mp
import numpy as np
from scipy.interpolate import CubicSpline


def smooth(u, t, d=12, lo=1e-4, hi=1e3, dps=80):
    """-> (x, dx/dt at the knots, lambda).  All linear algebra in `dps`-digit arithmetic: with the reference's default d = 12
    the entries of D'D span ~40 orders of magnitude, far beyond what a double-precision dense solve can be trusted with."""
    mp.mp.dps = dps
    n = len(t)
    d = min(d, n - 2)
    tt = [mp.mpf(float(v)) for v in t]
    # D by the textbook divided-difference recursion, row by row
    D = [[mp.mpf(1) if i == j else mp.mpf(0) for j in range(n)] for i in range(n)]
    for k in range(1, d + 1):
        D = [[k * (D[i + 1][j] - D[i][j]) / (tt[i + k] - tt[i]) for j in range(n)] for i in range(len(D) - 1)]
    Dm = mp.matrix(D)
    DtD = Dm.T * Dm
    E, Q = mp.eigsy(DtD)                      # D'D = Q diag(E) Q'
    uu = mp.matrix([mp.mpf(float(v)) for v in u])
    z = Q.T * uu
    E = [max(e, mp.mpf(0)) for e in E]
    z2 = [zi * zi for zi in z]

    def gcv(x):                               # x = log10(lambda); GCV = n |(I - H) u|^2 / tr(I - H)^2, H = (I + lambda^2 D'D)^-1
        l2 = mp.mpf(10) ** (2 * x)
        R = mp.mpf(0); T = mp.mpf(0)
        for a, b in zip(E, z2):
            om = l2 * a / (1 + l2 * a)
            R += om * om * b
            T += om
        return n * R / (T * T)

    xs = [mp.log10(lo) + (mp.log10(hi) - mp.log10(lo)) * q / 63 for q in range(64)]
    g = [gcv(x) for x in xs]
    k = min(range(64), key=lambda q: g[q])
    a, b = xs[max(k - 1, 0)], xs[min(k + 1, 63)]
    gr = (mp.sqrt(5) - 1) / 2
    c, dd = b - gr * (b - a), a + gr * (b - a)
    fc, fd = gcv(c), gcv(dd)
    for _ in range(160):
        if fc < fd:
            b, dd, fd = dd, c, fc
            c = b - gr * (b - a); fc = gcv(c)
        else:
            a, c, fc = c, dd, fd
            dd = a + gr * (b - a); fd = gcv(dd)
    xm = (a + b) / 2
    if gcv(xm) > g[k]:
        xm = xs[k]
    l2 = mp.mpf(10) ** (2 * xm)
    f = mp.matrix([1 / (1 + l2 * e) for e in E])
    x = Q * mp.matrix([f[i] * z[i] for i in range(n)])
    xv = np.array([float(v) for v in x])
    return xv, CubicSpline(np.asarray(t, dtype=np.float64), xv, bc_type="natural")(np.asarray(t, dtype=np.float64), 1), float(mp.mpf(10) ** xm)
