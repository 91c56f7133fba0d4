// The hot path: fused forward fixed-step RK + loss + discrete-adjoint reverse sweep, one sub-warp group of
// G lanes per trajectory segment.
//
// Mapping (DESIGN.md "Kernel"):
//   * lane g of a group owns hidden units j = g + G*i, i < HPL: its rows of W1, entries of b1 and columns of W2
//     live in REGISTERS for the whole kernel, and so do the matching gradient accumulators -- the MLP weight
//     gradient never leaves the register file until the single flush at the end of the persistent loop;
//   * the state vector, RK stage values, covariates and the adjoint lambda are replicated in every lane of the
//     group (d <= 8), so the only cross-lane traffic per RK stage is one all-reduce of DOUT partial sums
//     (y = W2 h) forward and one of D partial sums (W1^T zbar) backward -- reduce-scatter + all-gather shuffles;
//   * hidden activations h (one double per lane per stage) and the stage inputs are written to a per-warp
//     stash row (256 B, fully coalesced) on the forward sweep and read back in reverse, so tanh -- ~45% of the
//     FP64 issue slots -- is evaluated once per stage, never recomputed;
//   * a warp runs 32/G segments at once; segments in a warp that are shorter than the longest one keep
//     stepping with h = 0 and zero loss weight (no divergence around the shuffles);
//   * gradients of uhat columns go out with fire-and-forget FP64 atomics (<= 2 contributions per address).
#pragma once
#include "ude_device.cuh"
#include "ude_math.cuh"
#include "../../include/ude_cuda.h"
#include <type_traits>

namespace ude {
// piecewise-linear covariates with flat extrapolation (/root/reference/src/covariates.jl:16-19, :45-48).
// Each group caches the knot interval it is in (validity range, reference time, 1/dt, both knot values): stage times
// move monotonically inside a segment, so almost every evaluation is 4 FP64 instructions and no memory access; the
// refill (binary search over the series' knots) runs once per knot crossing.  Flat extrapolation is an interval with
// inv = 0 and equal end values.
template <int NX>
struct CovCache {
  static constexpr int N = NX > 0 ? NX : 1;
  double vlo[N], vhi[N], tref[N], inv[N], xlo[N], xhi[N];
};

template <int NX>
__device__ __forceinline__ void cov_reset(CovCache<NX>& c) {
#pragma unroll
  for (int v = 0; v < CovCache<NX>::N; ++v) { c.vlo[v] = 1.0; c.vhi[v] = 0.0; c.tref[v] = 0.0; c.inv[v] = 0.0; c.xlo[v] = 0.0; c.xhi[v] = 0.0; }
}

template <int NX>
__device__ __forceinline__ void cov_refill(const DevProblem& P, int series, int v, double t, CovCache<NX>& c) {
  const long long k0 = P.knot_off[(long long)series * NX + v];
  const int n = (int)(P.knot_off[(long long)series * NX + v + 1] - k0);
  const double* __restrict__ kt = P.knot_t + k0;
  const double* __restrict__ kx = P.knot_x + k0;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  if (n <= 0) { c.vlo[v] = -inf; c.vhi[v] = inf; c.tref[v] = 0.0; c.inv[v] = 0.0; c.xlo[v] = 0.0; c.xhi[v] = 0.0; return; }
  if (t < kt[0]) { c.vlo[v] = -inf; c.vhi[v] = kt[0]; c.tref[v] = kt[0]; c.inv[v] = 0.0; c.xlo[v] = kx[0]; c.xhi[v] = kx[0]; return; }
  if (t >= kt[n - 1]) { c.vlo[v] = kt[n - 1]; c.vhi[v] = inf; c.tref[v] = kt[n - 1]; c.inv[v] = 0.0; c.xlo[v] = kx[n - 1]; c.xhi[v] = kx[n - 1]; return; }
  int lo = 0, hi = n - 1;                  // kt[lo] <= t < kt[hi]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (kt[mid] <= t) lo = mid; else hi = mid; }
  c.vlo[v] = kt[lo]; c.vhi[v] = kt[lo + 1]; c.tref[v] = kt[lo]; c.inv[v] = P.knot_inv[k0 + lo];
  c.xlo[v] = kx[lo]; c.xhi[v] = kx[lo + 1];
}

template <int NX, class TX>
__device__ __forceinline__ void cov_eval(const DevProblem& P, int series, double t, CovCache<NX>& c,
                                         TX (&X)[NX > 0 ? NX : 1]) {
#pragma unroll
  for (int v = 0; v < NX; ++v) {
    if (!(t >= c.vlo[v] && t < c.vhi[v])) cov_refill<NX>(P, series, v, t, c);
    const double f = (t - c.tref[v]) * c.inv[v];
    X[v] = (TX)fma(f, c.xhi[v], (1.0 - f) * c.xlo[v]);
  }
}

// r' M r and (M + M') r for the d x d weight matrices of the state-space loss
template <int D>
__device__ __forceinline__ double quad_form(const double* __restrict__ M, const double* __restrict__ Msym,
                                            const double (&r)[D], double (&sym)[D]) {
  double q = 0.0;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double s = 0.0, ss = 0.0;
#pragma unroll
    for (int b = 0; b < D; ++b) { s = fma(M[a + D * b], r[b], s); ss = fma(Msym[a + D * b], r[b], ss); }
    q = fma(r[a], s, q);
    sym[a] = ss;
  }
  return q;
}

// ================================================================================================================
// LK selects the loss family at compile time (keeps each kernel's code inside the instruction cache):
//   LK_STATE  one observation interval per segment, target = the next uhat column (conditional likelihood / MSE loss)
//   LK_TRAJ   shooting / multiple-shooting block, targets = data columns
//   LK_DM     derivative matching: one RHS evaluation per column, no time stepping
enum { LK_STATE = 0, LK_TRAJ = 1, LK_DM = 2 };


namespace f64 {
typedef double real;
typedef double2 real2;
#include "ude_kernel_body.inc"
}  // namespace f64

namespace f32 {
typedef float real;
typedef float2 real2;
#include "ude_kernel_body.inc"
}  // namespace f32

using namespace f64;      // the FP64 names are the default (the tensor-core family and the registry helpers use them)

}  // namespace ude
