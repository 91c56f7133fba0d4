"""Derivative-matching pre-step: `RegularizationSmooth(u, t, d; alg = :gcv_svd)` + `DataInterpolations.derivative` at the
observation times (/root/reference/src/LossFunctionConstructors.jl:161-171, :510-522), batched on the GPU.

The smoother is third-party (DataInterpolations.jl `RegularizationSmooth` over RegularizationTools.jl, compat ranges in
/root/reference/Project.toml:47,65; source not under /root/reference), restated here from its published algorithm:
  * the curve values x at the data abscissae minimise  |x - u|^2 + lambda^2 |D x|^2  with D the d-th order divided-difference
    derivative matrix ((n-d) x n;  D_d = d diag(1 / (t[d+1:end] - t[1:end-d])) diff(D_{d-1}),  D_0 = I), unit weights;
    NOTE the reference passes its keyword `d` (default 12) as the DERIVATIVE ORDER;
  * lambda is chosen by generalised cross validation, GCV(lambda) = n |(I - H) u|^2 / tr(I - H)^2, over
    lambda in [1e-4, 1e3] (RegularizationTools.solve defaults), evaluated through the SVD of D (`:gcv_svd`);
  * the derivative is that of the NATURAL cubic spline through (t, x) (DataInterpolations.CubicSpline), taken at the knots.
Everything that depends on the time grid only -- the SVD of D and the spline-derivative matrix -- is built ONCE per distinct
grid on the host (NumPy, `grid_operators`); the per-row work (projection, the GCV minimisation, reconstruction, derivative) runs
in one CUDA kernel, one warp per (series, state) row (csrc/ude_smooth.cu, `ude_reg_smooth`).  No CPU fallback: without the
library / a GPU `reg_smooth_batch` raises.  oracle/smooth_oracle_np.py is the independent dense NumPy check (tests only).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

LAMBDA_MIN, LAMBDA_MAX = 1e-4, 1e3          # RegularizationTools.solve: lambda_1, lambda_2
N_SCAN = 64                                 # coarse GCV scan in log(lambda) before the root refinement


def derivative_matrix(t: np.ndarray, d: int) -> np.ndarray:
    """DataInterpolations._derivative_matrix: (n - d) x n matrix of the d-th divided-difference derivative."""
    n = t.shape[0]
    D = np.eye(n)
    for k in range(1, d + 1):
        dt = t[k:] - t[:-k]
        D = k * (np.diff(D, axis=0) / dt[:, None])
    return D


def spline_derivative_matrix(t: np.ndarray) -> np.ndarray:
    """n x n matrix M with (M x)[i] = s'(t_i), s the natural cubic spline through (t, x) (second derivatives z: the
    tridiagonal system of DataInterpolations.CubicSpline, z_1 = z_n = 0)."""
    n = t.shape[0]
    h = np.diff(t)
    if n == 2:
        return np.array([[-1.0, 1.0], [-1.0, 1.0]]) / h[0]
    A = np.zeros((n, n)); B = np.zeros((n, n))          # A z = B x
    A[0, 0] = A[-1, -1] = 1.0
    for i in range(1, n - 1):
        A[i, i - 1], A[i, i], A[i, i + 1] = h[i - 1], 2.0 * (h[i - 1] + h[i]), h[i]
        B[i, i - 1], B[i, i], B[i, i + 1] = 6.0 / h[i - 1], -6.0 / h[i - 1] - 6.0 / h[i], 6.0 / h[i]
    Z = np.linalg.solve(A, B)                            # z = Z x
    M = np.zeros((n, n))
    for i in range(n - 1):                               # left end of interval i
        M[i, i] -= 1.0 / h[i]; M[i, i + 1] += 1.0 / h[i]
        M[i] -= h[i] / 6.0 * (2.0 * Z[i] + Z[i + 1])
    i = n - 2                                            # right end of the last interval
    M[n - 1, i] -= 1.0 / h[i]; M[n - 1, i + 1] += 1.0 / h[i]
    M[n - 1] += h[i] / 6.0 * (Z[i] + 2.0 * Z[i + 1])
    return M


def grid_operators(t, d: int):
    """(Vt (r x n, rows = right singular vectors of D with non-zero singular value), s (r), M (n x n)) for one time grid."""
    t = np.asarray(t, dtype=np.float64)
    n = t.shape[0]
    d = int(min(d, n - 2))                               # a d-th derivative needs d + 1 points; keep at least one penalised direction
    if d < 1:
        raise ValueError("regularisation smoothing needs at least 3 points per series")
    D = derivative_matrix(t, d)
    # scale rows so the SVD is computed on O(1) numbers (d = 12 gives entries ~ dt^-12); the scale goes back into s
    scale = np.max(np.abs(D))
    _, s, Vt = np.linalg.svd(D / scale, full_matrices=False)
    # The null space of D is known exactly -- polynomials of degree < d on the grid -- and for high orders (the reference's
    # default d = 12) the double-precision SVD separates it from the penalised directions only to ~eps * cond(D) (1e-7 at
    # n = 60).  Build it from a Legendre basis (well conditioned) and re-orthonormalise the penalised directions against it.
    tau = 2.0 * (t - t[0]) / (t[-1] - t[0]) - 1.0
    Nq, _ = np.linalg.qr(np.polynomial.legendre.legvander(tau, d - 1))          # n x d, orthonormal
    W = Vt.T - Nq @ (Nq.T @ Vt.T)
    W = W - Nq @ (Nq.T @ W)                                                      # twice is enough
    Qw, Rw = np.linalg.qr(W)
    Vt = (Qw * np.sign(np.diag(Rw))).T                                           # keep the orientation of every direction
    return np.ascontiguousarray(Vt), np.ascontiguousarray(s * scale), np.ascontiguousarray(spline_derivative_matrix(t))


def reg_smooth_batch(U, t, d: int = 12, device: int = 0):
    """(smoothed, d/dt, lambda) for MANY rows U (rows x n) that share the time grid t, on the GPU."""
    from ._capi import UdeError, load_library
    U = np.ascontiguousarray(U, dtype=np.float64)
    rows, n = U.shape
    Vt, s, M = grid_operators(t, d)
    out_u = np.empty_like(U); out_du = np.empty_like(U); lam = np.empty(rows)
    lib = load_library()
    rc = lib.ude_reg_smooth(int(device), int(n), int(s.shape[0]), int(rows), Vt.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                            M.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), float(LAMBDA_MIN), float(LAMBDA_MAX),
                            out_u.ctypes.data_as(C.c_void_p), out_du.ctypes.data_as(C.c_void_p), lam.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise UdeError(rc, lib.ude_last_error(None).decode())
    return out_u, out_du, lam
