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
#include "ude_rhs.cuh"
#include <type_traits>

namespace ude {

template <class R, int HPL>
struct LaneNet {
  double W1[HPL][R::DIN];
  double b1[HPL];
  double W2[R::DOUT][HPL];
  double b2[R::DOUT];
};
template <class R, int HPL>
struct LaneGrad {
  double W1[HPL][R::DIN];
  double b1[HPL];
  double W2[R::DOUT][HPL];
  double b2[R::DOUT];
  double sc[R::NS > 0 ? R::NS : 1];
};

template <class R, int G, int HPL>
struct Geo {
  static constexpr int D = R::D, NX = R::NX, DIN = R::DIN, DOUT = R::DOUT, NS = R::NS;
  static constexpr int NXA = NX > 0 ? NX : 1;
  static constexpr int NU = D + NX + (R::NEED_Y ? DOUT : 0);   // uniform doubles stashed per stage
  static constexpr int ROWS = HPL + (NU + G - 1) / G;          // 256-byte stash rows per stage
  static constexpr int SPW = 32 / G;                            // segments per warp
  // shared-memory weight record of one lane (WS kernels): per owned hidden unit [W1 row | b1 | pad][W2 column | pad],
  // both padded to an even count so every access is one LDS.128; the record stride is an odd number of 16-byte
  // units, which spreads the G records of a warp over all banks (lanes of equal g broadcast).
  static constexpr int DINP = (DIN + 1 + 1) / 2 * 2;            // inputs + the constant-1 bias input, even
  static constexpr int DOUTP = (DOUT + 1) / 2 * 2;
  static constexpr int HB2 = (DINP + DOUTP) / 2;                // double2 per hidden unit
  static constexpr int WREC2 = (HPL * HB2) | 1;                 // double2 per lane record (odd)
};

template <class R, int HPL>
struct LaneNetS {            // weights in shared memory, only the uniform output bias in registers
  const double2* rec;
  double b2[R::DOUT];
};

template <class R, int G, int HPL>
__device__ __forceinline__ void load_net(const DevProblem& P, const double* __restrict__ pm, int g,
                                         LaneNet<R, HPL>& n, RhsCtx& ctx) {
#pragma unroll
  for (int i = 0; i < HPL; ++i) {
    const int j = g + G * i;
    const bool ok = j < P.H;
    const int jj = ok ? j : 0;
#pragma unroll
    for (int k = 0; k < R::DIN; ++k) { double v = pm[P.off_w1 + jj + (long long)P.H * k]; n.W1[i][k] = ok ? v : 0.0; }
    { double v = pm[P.off_b1 + jj]; n.b1[i] = ok ? v : 0.0; }
#pragma unroll
    for (int o = 0; o < R::DOUT; ++o) { double v = pm[P.off_w2 + o + (long long)P.nn_out * jj]; n.W2[o][i] = ok ? v : 0.0; }
  }
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) n.b2[o] = pm[P.off_b2 + o];
#pragma unroll
  for (int k = 0; k < R::NS; ++k) ctx.sc[k] = pm[P.off_scalar[k]];
  ctx.time_scale = P.time_scale; ctx.time_shift = P.time_shift; ctx.series_param = 0.0;
}

template <class R, int HPL>
__device__ __forceinline__ void zero_grad(LaneGrad<R, HPL>& gr) {
#pragma unroll
  for (int i = 0; i < HPL; ++i) {
#pragma unroll
    for (int k = 0; k < R::DIN; ++k) gr.W1[i][k] = 0.0;
    gr.b1[i] = 0.0;
#pragma unroll
    for (int o = 0; o < R::DOUT; ++o) gr.W2[o][i] = 0.0;
  }
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) gr.b2[o] = 0.0;
#pragma unroll
  for (int k = 0; k < (R::NS > 0 ? R::NS : 1); ++k) gr.sc[k] = 0.0;
}

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

template <int NX>
__device__ __forceinline__ void cov_eval(const DevProblem& P, int series, double t, CovCache<NX>& c,
                                         double (&X)[NX > 0 ? NX : 1]) {
#pragma unroll
  for (int v = 0; v < NX; ++v) {
    if (!(t >= c.vlo[v] && t < c.vhi[v])) cov_refill<NX>(P, series, v, t, c);
    const double f = (t - c.tref[v]) * c.inv[v];
    X[v] = fma(f, c.xhi[v], (1.0 - f) * c.xlo[v]);
  }
}

// ---- one RHS evaluation: y = W2 tanh(W1 x + b1) + b2, du = combine(U, X, y) ------------------------------
template <class R, int G, int HPL>
__device__ __forceinline__ void stage_forward(const LaneNet<R, HPL>& n, const RhsCtx& ctx, const double* s_tab, int lane,
                                              const double (&U)[R::D], const double* X, double t,
                                              double (&du)[R::D], double (&hh)[HPL], double (&y)[R::DOUT]) {
  double x[R::DIN];
  R::input(U, X, t, ctx, x);
  double z[HPL];
#pragma unroll
  for (int i = 0; i < HPL; ++i) z[i] = n.b1[i];
#pragma unroll
  for (int k = 0; k < R::DIN; ++k) {
#pragma unroll
    for (int i = 0; i < HPL; ++i) z[i] = fma(n.W1[i][k], x[k], z[i]);
  }
  ude_tanh_tab_vec<HPL>(z, hh, s_tab);
  double part[R::DOUT];
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) {
    double p = n.W2[o][0] * hh[0];
#pragma unroll
    for (int i = 1; i < HPL; ++i) p = fma(n.W2[o][i], hh[i], p);
    part[o] = p;
  }
  group_allreduce<G, R::DOUT>(part, lane);
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) y[o] = part[o] + n.b2[o];
  R::combine(U, X, y, ctx, du);
}

// ---- its vector-Jacobian product: ub += (df/dU)^T kb, gradient accumulators += (df/dtheta)^T kb ------------
template <class R, int G, int HPL>
__device__ __forceinline__ void stage_backward(const LaneNet<R, HPL>& n, LaneGrad<R, HPL>& gr, const RhsCtx& ctx, int lane,
                                               const double (&U)[R::D], const double* X, double t,
                                               const double (&hh)[HPL], const double (&y)[R::DOUT],
                                               const double (&kb)[R::D], double (&ub)[R::D], double& gseries) {
  double x[R::DIN];
  R::input(U, X, t, ctx, x);
  double yb[R::DOUT];
  R::combine_vjp(U, X, y, ctx, kb, yb, ub, gr.sc, gseries);
  double xpart[R::D];
#pragma unroll
  for (int k = 0; k < R::D; ++k) xpart[k] = 0.0;
#pragma unroll
  for (int i = 0; i < HPL; ++i) {
    double hb = n.W2[0][i] * yb[0];
#pragma unroll
    for (int o = 1; o < R::DOUT; ++o) hb = fma(n.W2[o][i], yb[o], hb);
    const double zb = hb * fma(-hh[i], hh[i], 1.0);
#pragma unroll
    for (int o = 0; o < R::DOUT; ++o) gr.W2[o][i] = fma(yb[o], hh[i], gr.W2[o][i]);
    gr.b1[i] += zb;
#pragma unroll
    for (int k = 0; k < R::DIN; ++k) gr.W1[i][k] = fma(zb, x[k], gr.W1[i][k]);
#pragma unroll
    for (int k = 0; k < R::D; ++k) xpart[k] = fma(n.W1[i][k], zb, xpart[k]);
  }
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) gr.b2[o] += yb[o];
  group_allreduce<G, R::D>(xpart, lane);
#pragma unroll
  for (int k = 0; k < R::D; ++k) ub[k] += xpart[k];
}

// ---- the same two routines with the lane's weights read from shared memory (LDS.128, two weights per load) -------
// Used when the register-resident copy would push the kernel past the register file (H >= 32): the gradient
// accumulators stay in registers, the weights are re-read each stage on the otherwise idle LSU pipe.
template <class R, int G, int HPL>
__device__ __forceinline__ void stage_forward(const LaneNetS<R, HPL>& n, const RhsCtx& ctx, const double* s_tab, int lane,
                                              const double (&U)[R::D], const double* X, double t,
                                              double (&du)[R::D], double (&hh)[HPL], double (&y)[R::DOUT]) {
  using GG = Geo<R, G, HPL>;
  double x[R::DIN];
  R::input(U, X, t, ctx, x);
  double xx[GG::DINP];
#pragma unroll
  for (int k = 0; k < GG::DINP; ++k) xx[k] = (k < R::DIN) ? x[k < R::DIN ? k : 0] : (k == R::DIN ? 1.0 : 0.0);
  double z[HPL];
#pragma unroll
  for (int i = 0; i < HPL; ++i) z[i] = 0.0;
#pragma unroll
  for (int kp = 0; kp < GG::DINP / 2; ++kp) {
#pragma unroll
    for (int i = 0; i < HPL; ++i) {
      const double2 w = n.rec[i * GG::HB2 + kp];
      z[i] = fma(w.x, xx[2 * kp], z[i]);
      z[i] = fma(w.y, xx[2 * kp + 1], z[i]);
    }
  }
  ude_tanh_tab_vec<HPL>(z, hh, s_tab);
  double part[GG::DOUTP];
#pragma unroll
  for (int o = 0; o < GG::DOUTP; ++o) part[o] = 0.0;
#pragma unroll
  for (int i = 0; i < HPL; ++i) {
#pragma unroll
    for (int op = 0; op < GG::DOUTP / 2; ++op) {
      const double2 w = n.rec[i * GG::HB2 + GG::DINP / 2 + op];
      part[2 * op] = fma(w.x, hh[i], part[2 * op]);
      part[2 * op + 1] = fma(w.y, hh[i], part[2 * op + 1]);
    }
  }
  double pr[R::DOUT];
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) pr[o] = part[o];
  group_allreduce<G, R::DOUT>(pr, lane);
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) y[o] = pr[o] + n.b2[o];
  R::combine(U, X, y, ctx, du);
}

template <class R, int G, int HPL>
__device__ __forceinline__ void stage_backward(const LaneNetS<R, HPL>& n, LaneGrad<R, HPL>& gr, const RhsCtx& ctx, int lane,
                                               const double (&U)[R::D], const double* X, double t,
                                               const double (&hh)[HPL], const double (&y)[R::DOUT],
                                               const double (&kb)[R::D], double (&ub)[R::D], double& gseries) {
  using GG = Geo<R, G, HPL>;
  double x[R::DIN];
  R::input(U, X, t, ctx, x);
  double yb[R::DOUT];
  R::combine_vjp(U, X, y, ctx, kb, yb, ub, gr.sc, gseries);
  double ybp[GG::DOUTP];
#pragma unroll
  for (int o = 0; o < GG::DOUTP; ++o) ybp[o] = (o < R::DOUT) ? yb[o < R::DOUT ? o : 0] : 0.0;
  double hb[HPL], zb[HPL];
#pragma unroll
  for (int i = 0; i < HPL; ++i) hb[i] = 0.0;
#pragma unroll
  for (int op = 0; op < GG::DOUTP / 2; ++op) {
#pragma unroll
    for (int i = 0; i < HPL; ++i) {
      const double2 w = n.rec[i * GG::HB2 + GG::DINP / 2 + op];
      hb[i] = fma(w.x, ybp[2 * op], hb[i]);
      hb[i] = fma(w.y, ybp[2 * op + 1], hb[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < HPL; ++i) zb[i] = hb[i] * fma(-hh[i], hh[i], 1.0);
  double xpart[(R::D + 1) / 2 * 2];
#pragma unroll
  for (int k = 0; k < (R::D + 1) / 2 * 2; ++k) xpart[k] = 0.0;
#pragma unroll
  for (int kp = 0; kp < (R::D + 1) / 2; ++kp) {
#pragma unroll
    for (int i = 0; i < HPL; ++i) {
      const double2 w = n.rec[i * GG::HB2 + kp];
      xpart[2 * kp] = fma(w.x, zb[i], xpart[2 * kp]);
      xpart[2 * kp + 1] = fma(w.y, zb[i], xpart[2 * kp + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < HPL; ++i) {
#pragma unroll
    for (int o = 0; o < R::DOUT; ++o) gr.W2[o][i] = fma(yb[o], hh[i], gr.W2[o][i]);
    gr.b1[i] += zb[i];
#pragma unroll
    for (int k = 0; k < R::DIN; ++k) gr.W1[i][k] = fma(zb[i], x[k], gr.W1[i][k]);
  }
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) gr.b2[o] += yb[o];
  double xr[R::D];
#pragma unroll
  for (int k = 0; k < R::D; ++k) xr[k] = xpart[k];
  group_allreduce<G, R::D>(xr, lane);
#pragma unroll
  for (int k = 0; k < R::D; ++k) ub[k] += xr[k];
}

// fill this warp's shared-memory weight records (one per g) and the uniform leaves
template <class R, int G, int HPL>
__device__ __forceinline__ void load_net(const DevProblem& P, const double* __restrict__ pm, int lane, int g,
                                         double* s_w, LaneNetS<R, HPL>& n, RhsCtx& ctx) {
  using GG = Geo<R, G, HPL>;
  __syncwarp();
  double* rec = s_w + (size_t)g * GG::WREC2 * 2;
  if (lane < G) {
#pragma unroll
    for (int i = 0; i < HPL; ++i) {
      const int j = g + G * i;
      const bool ok = j < P.H;
      const int jj = ok ? j : 0;
#pragma unroll
      for (int k = 0; k < GG::DINP; ++k) {
        double v = 0.0;
        if (k < R::DIN) v = pm[P.off_w1 + jj + (long long)P.H * k];
        else if (k == R::DIN) v = pm[P.off_b1 + jj];
        rec[i * GG::HB2 * 2 + k] = ok ? v : 0.0;
      }
#pragma unroll
      for (int o = 0; o < GG::DOUTP; ++o) {
        double v = 0.0;
        if (o < R::DOUT) v = pm[P.off_w2 + o + (long long)P.nn_out * jj];
        rec[i * GG::HB2 * 2 + GG::DINP + o] = ok ? v : 0.0;
      }
    }
  }
  __syncwarp();
  n.rec = reinterpret_cast<const double2*>(rec);
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) n.b2[o] = pm[P.off_b2 + o];
#pragma unroll
  for (int k = 0; k < R::NS; ++k) ctx.sc[k] = pm[P.off_scalar[k]];
  ctx.time_scale = P.time_scale; ctx.time_shift = P.time_shift; ctx.series_param = 0.0;
}

// ---- stash rows: [HPL rows of h | ceil(NU/G) rows holding the uniform values spread over the group's lanes] --
template <class R, int G, int HPL>
__device__ __forceinline__ void stash_store(double* __restrict__ rec, int lane, int g, const double (&hh)[HPL],
                                            const double (&U)[R::D], const double* X, const double (&y)[R::DOUT]) {
  using GG = Geo<R, G, HPL>;
#pragma unroll
  for (int i = 0; i < HPL; ++i) rec[i * 32 + lane] = hh[i];
#pragma unroll
  for (int e = 0; e < GG::NU; ++e) {
    const double v = (e < R::D) ? U[e < R::D ? e : 0]
                   : (e < R::D + R::NX) ? X[(e >= R::D && e < R::D + R::NX) ? e - R::D : 0]
                   : y[(e >= R::D + R::NX) ? e - R::D - R::NX : 0];
    if (g == e % G) rec[(HPL + e / G) * 32 + lane] = v;
  }
}
template <class R, int G, int HPL>
__device__ __forceinline__ void stash_load(const double* __restrict__ rec, int lane, int g, double (&hh)[HPL],
                                           double (&U)[R::D], double* X, double (&y)[R::DOUT]) {
  using GG = Geo<R, G, HPL>;
  const int gbase = lane - g;
#pragma unroll
  for (int i = 0; i < HPL; ++i) hh[i] = rec[i * 32 + lane];
#pragma unroll
  for (int e = 0; e < GG::NU; ++e) {
    const double v = rec[(HPL + e / G) * 32 + gbase + e % G];
    if (e < R::D) U[e < R::D ? e : 0] = v;
    else if (e < R::D + R::NX) X[(e >= R::D && e < R::D + R::NX) ? e - R::D : 0] = v;
    else y[(e >= R::D + R::NX) ? e - R::D - R::NX : 0] = v;
  }
}

// ---- one explicit RK step forward; GRAD: record every stage -------------------------------------------------
template <class R, int G, int HPL, int SOLVER, bool GRAD, class NetT>
__device__ __forceinline__ void rk_forward(const DevProblem& P, const NetT& n, const RhsCtx& ctx, const double* s_tab,
                                           int lane, int g, int series, CovCache<R::NX>& cur,
                                           double (&u)[R::D], double t, double hs, double* __restrict__ step_rec) {
  using T = Tab<SOLVER>;
  using GG = Geo<R, G, HPL>;
  constexpr int D = R::D;
  double k[T::S][D];
  double acc[D];
#pragma unroll
  for (int j = 0; j < D; ++j) acc[j] = 0.0;
#pragma unroll
  for (int i = 0; i < T::S; ++i) {
    double U[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double a = 0.0;
#pragma unroll
      for (int l = 0; l < i; ++l) if (T::a(i, l) != 0.0) a = fma(T::a(i, l), k[l][j], a);
      U[j] = (i == 0) ? u[j] : fma(hs, a, u[j]);
    }
    const double ts = t + T::c(i) * hs;
    double X[GG::NXA];
    X[0] = 0.0;
    if (R::NX > 0) cov_eval<R::NX>(P, series, ts, cur, X);
    double du[D], hh[HPL], y[R::DOUT];
    stage_forward<R, G, HPL>(n, ctx, s_tab, lane, U, X, ts, du, hh, y);
    if (GRAD) stash_store<R, G, HPL>(step_rec + i * GG::ROWS * 32, lane, g, hh, U, X, y);
#pragma unroll
    for (int j = 0; j < D; ++j) { k[i][j] = du[j]; acc[j] = fma(T::b(i), du[j], acc[j]); }
  }
#pragma unroll
  for (int j = 0; j < D; ++j) u[j] = fma(hs, acc[j], u[j]);
}

// ---- its discrete adjoint; returns the step's initial state in U0 ---------------------------
// step_rec points at the step's record in the warp's shared-memory staging buffer (filled by a TMA bulk load).
template <class R, int G, int HPL, int SOLVER, class NetT>
__device__ __forceinline__ void rk_reverse(const NetT& n, LaneGrad<R, HPL>& gr, const RhsCtx& ctx, int lane, int g,
                                           double (&lam)[R::D], double t, double hs, const double* __restrict__ step_rec,
                                           double (&U0)[R::D], double& gseries) {
  using T = Tab<SOLVER>;
  using GG = Geo<R, G, HPL>;
  constexpr int D = R::D;
  double Ub[T::S][D];
  double lamacc[D];
#pragma unroll
  for (int j = 0; j < D; ++j) lamacc[j] = 0.0;
#pragma unroll
  for (int ii = 0; ii < T::S; ++ii) {
    const int i = T::S - 1 - ii;
    double hh[HPL], U[D], X[GG::NXA], y[R::DOUT];
    X[0] = 0.0;
#pragma unroll
    for (int o = 0; o < R::DOUT; ++o) y[o] = 0.0;
    stash_load<R, G, HPL>(step_rec + i * GG::ROWS * 32, lane, g, hh, U, X, y);
    double kb[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double a = T::b(i) * lam[j];
#pragma unroll
      for (int l = i + 1; l < T::S; ++l) if (T::a(l, i) != 0.0) a = fma(T::a(l, i), Ub[l][j], a);
      kb[j] = hs * a;
    }
    const double ts = t + T::c(i) * hs;
    double ub[D];
#pragma unroll
    for (int j = 0; j < D; ++j) ub[j] = 0.0;
    stage_backward<R, G, HPL>(n, gr, ctx, lane, U, X, ts, hh, y, kb, ub, gseries);
#pragma unroll
    for (int j = 0; j < D; ++j) { Ub[i][j] = ub[j]; lamacc[j] += ub[j]; }
    if (i == 0) {
#pragma unroll
      for (int j = 0; j < D; ++j) U0[j] = U[j];
    }
  }
#pragma unroll
  for (int j = 0; j < D; ++j) lam[j] += lamacc[j];
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

// sum over the groups of a warp (lanes with equal g), result valid in every lane
template <int G>
__device__ __forceinline__ double across_groups(double v) {
#pragma unroll
  for (int off = G; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// ---- flush the register-resident gradient of one model -------------------------------------------------------
// Sum over the warp's groups with xor shuffles, then hand every (offset, value) pair to `put` from exactly one lane.
template <class R, int G, int HPL, class Put>
__device__ __forceinline__ void emit_grad(const DevProblem& P, const LaneGrad<R, HPL>& gr, int lane, int g, Put put) {
#pragma unroll
  for (int i = 0; i < HPL; ++i) {
    const int j = g + G * i;
    const bool w = (lane < G) && (j < P.H);
#pragma unroll
    for (int k = 0; k < R::DIN; ++k) {
      const double v = across_groups<G>(gr.W1[i][k]);
      if (w) put(P.off_w1 + j + (long long)P.H * k, v);
    }
    {
      const double v = across_groups<G>(gr.b1[i]);
      if (w) put(P.off_b1 + j, v);
    }
#pragma unroll
    for (int o = 0; o < R::DOUT; ++o) {
      const double v = across_groups<G>(gr.W2[o][i]);
      if (w) put(P.off_w2 + o + (long long)P.nn_out * j, v);
    }
  }
#pragma unroll
  for (int o = 0; o < R::DOUT; ++o) {
    const double v = across_groups<G>(gr.b2[o]);
    if (lane == 0) put(P.off_b2 + o, v);
  }
#pragma unroll
  for (int k = 0; k < R::NS; ++k) {
    const double v = across_groups<G>(gr.sc[k]);
    if (lane == 0) put(P.off_scalar[k], v);
  }
}

// many independent models (CV folds x ensemble members): FP64 atomics straight into that model's gradient
template <class R, int G, int HPL, bool GRAD>
__device__ __forceinline__ void flush_model(const DevProblem& P, const LaneGrad<R, HPL>& gr, double loss, int lane, int g, int model) {
  if (GRAD) {
    double* dst = P.grad + P.theta_base[model] + (long long)R::D * (P.model_col0[model + 1] - P.model_col0[model]);
    emit_grad<R, G, HPL>(P, gr, lane, g, [&](long long idx, double v) { atomicAdd(dst + idx, v); });
  }
  const double lsum = across_groups<G>(loss);
  if (lane == 0) atomicAdd(P.model_loss + model, lsum);
}

// one shared model: warps of the block add into shared memory one after the other (fixed order), the block
// writes one partial row; ude_finalize sums the rows in a fixed order => bitwise reproducible gradients.
template <class R, int G, int HPL, bool GRAD>
__device__ __forceinline__ void flush_block(const DevProblem& P, const LaneGrad<R, HPL>& gr, double loss, int lane, int g,
                                            double* s_red) {
  const int nw = blockDim.x >> 5, wib = threadIdx.x >> 5;
  for (int w = 0; w < nw; ++w) {
    if (wib == w) {
      if (GRAD) emit_grad<R, G, HPL>(P, gr, lane, g, [&](long long idx, double v) { s_red[idx] += v; });
      const double lsum = across_groups<G>(loss);
      if (lane == 0) s_red[P.n_pm] += lsum;
    }
    __syncthreads();
  }
  double* row = P.block_part + (long long)blockIdx.x * (P.n_pm + 1);
  for (long long k = threadIdx.x; k <= P.n_pm; k += blockDim.x) row[k] = s_red[k];
}

// ================================================================================================================
// LK selects the loss family at compile time (keeps each kernel's code inside the instruction cache):
//   LK_STATE  one observation interval per segment, target = the next uhat column (conditional likelihood / MSE loss)
//   LK_TRAJ   shooting / multiple-shooting block, targets = data columns
//   LK_DM     derivative matching: one RHS evaluation per column, no time stepping
enum { LK_STATE = 0, LK_TRAJ = 1, LK_DM = 2 };

template <class R, int G, int HPL, int SOLVER, int LK, bool GRAD, int MINB, bool WS>
__global__ void __launch_bounds__(kBlockThreads, MINB) ude_seg_kernel(const DevProblem P) {
  using GG = Geo<R, G, HPL>;
  using T = Tab<SOLVER>;
  constexpr int D = R::D;
  __shared__ double s_tab[64];
  __shared__ __align__(8) uint64_t s_bar[kBlockThreads / 32][2];
  // dynamic shared memory: [n_pm + 1 doubles (block-level gradient/loss partial, single model), padded to 16 B]
  //                        [per warp: two step records of S * ROWS rows of 256 B -- the TMA staging ring of the stash]
  extern __shared__ __align__(16) double s_dyn[];
  double* s_red = s_dyn;
  ude_load_exp2_tab(s_tab);
  if (P.n_models == 1) for (long long k = threadIdx.x; k <= P.n_pm; k += blockDim.x) s_red[k] = 0.0;
  constexpr int kStepDoubles = T::S * GG::ROWS * 32;
  constexpr uint32_t kStepBytes = kStepDoubles * 8;
  constexpr int kWDoubles = WS ? G * GG::WREC2 * 2 : 0;      // this warp's weight records
  constexpr int kStageDoubles = (GRAD && LK != LK_DM) ? 2 * kStepDoubles : 0;
  double* s_warp = s_dyn + ((P.n_models == 1 ? P.n_pm + 1 : 0) + 1) / 2 * 2 + (size_t)(threadIdx.x >> 5) * (kWDoubles + kStageDoubles);
  double* s_w = s_warp;
  double* s_stage = s_warp + kWDoubles;
  uint64_t* bar = s_bar[threadIdx.x >> 5];
  if (GRAD && LK != LK_DM && (threadIdx.x & 31) == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  uint32_t bar_phase = 0;      // bit b = parity the next wait on bar[b] expects
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int g = lane % G, q = lane / G;
  const long long warps_per_block = blockDim.x >> 5;
  const long long warp_global = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const long long total_warps = (long long)gridDim.x * warps_per_block;
  double* __restrict__ stash = GRAD ? P.stash + warp_global * P.stash_rows_per_warp * 32 : nullptr;
  constexpr long long rec_stride = kStepDoubles;
  const int S = P.substeps;

  typename std::conditional<WS, LaneNetS<R, HPL>, LaneNet<R, HPL>>::type net;
  LaneGrad<R, HPL> gr;
  RhsCtx ctx;
  zero_grad<R, HPL>(gr);
  double loss = 0.0;
  int cur_model = -1;

  for (long long task = P.task_begin + warp_global; task < P.n_tasks; task += total_warps) {
    const Seg sg = P.segs[task * GG::SPW + q];
    const int model = __shfl_sync(0xffffffffu, sg.model, 0);
    const long long ncols_m = P.model_col0[model + 1] - P.model_col0[model];
    const double* __restrict__ pm = P.theta + P.theta_base[model] + (long long)D * ncols_m;
    if (model != cur_model) {
      if (cur_model >= 0) {
        flush_model<R, G, HPL, GRAD>(P, gr, loss, lane, g, cur_model);
        zero_grad<R, HPL>(gr);
        loss = 0.0;
      }
      if constexpr (WS) load_net<R, G, HPL>(P, pm, lane, g, s_w, net, ctx);
      else load_net<R, G, HPL>(P, pm, g, net, ctx);
      cur_model = model;
    }
    const bool is_null = (sg.flags & SEG_NULL) != 0;
    const long long uoff = P.theta_base[model] - (long long)D * P.model_col0[model];   // index uhat with GLOBAL columns
    const double* __restrict__ uh = P.theta + uoff;
    double* __restrict__ guh = GRAD ? P.grad + uoff : nullptr;
    const int series = sg.series;
    const long long series_local = series - P.model_series0[model];
    if (R::SERIES) ctx.series_param = pm[P.off_series_param + series_local];
    double gseries = 0.0;
    CovCache<R::NX> cur;
    cov_reset<R::NX>(cur);

    if (LK == LK_DM) {
      // ---- one column of the derivative-matching loss: LFC.jl:181-186, :525-533 ----------------------------------
      const long long c = sg.col;
      double U[D], X[GG::NXA], du[D], hh[HPL], y[R::DOUT], kb[D], ub[D];
      X[0] = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) U[j] = P.dm_u[D * c + j];
      const double ts = P.times[c];
      if (R::NX > 0) cov_eval<R::NX>(P, series, ts, cur, X);
      stage_forward<R, G, HPL>(net, ctx, s_tab, lane, U, X, ts, du, hh, y);
      const double w = is_null ? 0.0 : 1.0;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const double r = P.dm_dudt[D * c + j] - du[j];
        loss = fma(w * r, r, loss);
        kb[j] = -2.0 * w * r; ub[j] = 0.0;
      }
      if (GRAD) stage_backward<R, G, HPL>(net, gr, ctx, lane, U, X, ts, hh, y, kb, ub, gseries);
    } else {
      // ---- a segment = [pre-roll step] + maxL intervals of S steps; groups with fewer intervals step with h = 0 ----
      const bool multi = P.loss_kind == UDE_LOSS_MULTI_SHOOT;
      const bool skip = is_null || (LK == LK_STATE && (sg.flags & SEG_SKIP_PROC));
      long long c0;            // first column of the segment
      int L, maxL, pre, first_scored;
      bool from_theta, has_prev = true;
      double wpt = 0.0, eps = 0.0;
      if (LK == LK_STATE) {
        const long long c = sg.col;
        has_prev = !is_null && (c - 1 >= P.starts[series]);
        c0 = has_prev ? c - 1 : c;
        L = (skip || !has_prev) ? 0 : 1; maxL = 1; pre = 0; first_scored = 0; from_theta = true;
      } else {
        c0 = sg.col;
        L = is_null ? 0 : sg.L;
        maxL = (int)__reduce_max_sync(0xffffffffu, (unsigned)L);
        from_theta = multi || P.shoot_from_theta;
        first_scored = multi ? 0 : P.shoot_first_scored;
        wpt = is_null ? 0.0 : (multi ? 1.0 / ((double)D * (double)P.lengths[series]) : P.shoot_weight[model]);
        eps = multi ? P.preroll_eps : 0.0;
        pre = eps > 0.0 ? 1 : 0;
      }
      const int nsteps = pre + maxL * S;
      double u[D];
#pragma unroll
      for (int j = 0; j < D; ++j) u[j] = from_theta ? uh[D * c0 + j] : P.data[D * c0 + j];

      // ---------------- forward sweep: ONE inlined RK step in a flat loop over the segment's steps ----------------
      {
        int i = pre ? -1 : 0, n = pre ? S - 1 : 0;
        double t0 = P.times[c0] - eps, hs = is_null ? 0.0 : eps, nd = 0.0;
        for (int k = 0; k < nsteps; ++k) {
          if (k >= pre && n == 0) {
            // entering interval i at observation point i
            const int ic = min(i, L);
            if (LK == LK_TRAJ) {
              const double w = (i <= L && i >= first_scored) ? wpt : 0.0;
#pragma unroll
              for (int j = 0; j < D; ++j) { const double r = P.data[D * (c0 + ic) + j] - u[j]; loss = fma(w * r, r, loss); }
              if (!GRAD && P.traj_out != nullptr && i <= L && !is_null) {
#pragma unroll
                for (int j = 0; j < D; ++j) if (g == j % G) P.traj_out[D * (c0 + i) + j] = u[j];
              }
            }
            t0 = P.times[c0 + ic];
            hs = (i < L) ? (P.times[c0 + ic + 1] - t0) / S : 0.0;
            nd = 0.0;
          }
          double* sbuf = s_stage + (k & 1) * kStepDoubles;
          if (GRAD) {   // the bulk store that last read this buffer (two steps ago) must have drained it
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
          rk_forward<R, G, HPL, SOLVER, GRAD>(P, net, ctx, s_tab, lane, g, series, cur, u, fma(nd, hs, t0), hs, sbuf);
          if (GRAD) {   // hand the step record to the TMA: smem -> this warp's stash slot k
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_store(stash + (long long)k * rec_stride, sbuf, kStepBytes); bulk_commit(); }
          }
          nd += 1.0;
          if (++n == S) { n = 0; ++i; }
        }
        if (GRAD) {     // stash complete and visible before the reverse sweep's bulk loads; start fetching the last step
          if (lane == 0) {
            bulk_wait_all<0>();
            if (nsteps > 0) {
              const int b = (nsteps - 1) & 1;
              mbar_expect_tx(&bar[b], kStepBytes);
              bulk_load(s_stage + b * kStepDoubles, stash + (long long)(nsteps - 1) * rec_stride, kStepBytes, &bar[b]);
            }
          }
          __syncwarp();
        }
      }
      // ---------------- loss at the last point, seed of the adjoint ---------------------------------------------------
      double lam[D];
      double gc[D], gp[D];
      if (LK == LK_STATE) {
        // LFC.jl:38-44 / helpers.jl:31-37: (uhat_c - F(uhat_{c-1}))' B (.) weighted; observation terms LFC.jl:30-34
        const long long c = sg.col;
        const double dt = P.times[c] - P.times[c0];
        double r[D], sym[D];
#pragma unroll
        for (int j = 0; j < D; ++j) r[j] = uh[D * c + j] - u[j];
        const double wsc = (L == 1) ? P.proc_scale[model] * (P.proc_inv_dt ? 1.0 / dt : 1.0) : 0.0;
        loss = fma(wsc, quad_form<D>(P.B, P.Bsym, r, sym), loss);
#pragma unroll
        for (int j = 0; j < D; ++j) { lam[j] = -wsc * sym[j]; gc[j] = wsc * sym[j]; gp[j] = 0.0; }
        const double os = is_null ? 0.0 : P.obs_scale[model];
        {
          double e[D], es[D];
#pragma unroll
          for (int j = 0; j < D; ++j) e[j] = P.data[D * c + j] - uh[D * c + j];
          loss = fma(os, quad_form<D>(P.A, P.Asym, e, es), loss);
#pragma unroll
          for (int j = 0; j < D; ++j) gc[j] = fma(-os, es[j], gc[j]);
        }
        if ((sg.flags & SEG_OBS_FIRST) && has_prev) {
          double e[D], es[D];
#pragma unroll
          for (int j = 0; j < D; ++j) e[j] = P.data[D * c0 + j] - uh[D * c0 + j];
          loss = fma(os, quad_form<D>(P.A, P.Asym, e, es), loss);
#pragma unroll
          for (int j = 0; j < D; ++j) gp[j] = -os * es[j];
        }
        if (!GRAD && P.traj_out != nullptr && !is_null) {
#pragma unroll
          for (int j = 0; j < D; ++j) if (g == j % G) {
            if (L == 1) P.traj_out[D * c + j] = u[j];
            if (!has_prev || (sg.flags & SEG_OBS_FIRST)) P.traj_out[D * c0 + j] = uh[D * c0 + j];
          }
        }
      } else {
        const double w = (maxL <= L && maxL >= first_scored) ? wpt : 0.0;
        const int ic = min(maxL, L);
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const double r = P.data[D * (c0 + ic) + j] - u[j];
          loss = fma(w * r, r, loss);
          lam[j] = -2.0 * w * r;
        }
        // the block's end point is written by the block that STARTS there when one exists (a later block overwrites
        // it in the reference, LFC.jl:336, :751): full blocks of a multiple-shooting series leave it alone
        if (!GRAD && P.traj_out != nullptr && maxL <= L && !is_null && !(multi && L == P.pred_length)) {
#pragma unroll
          for (int j = 0; j < D; ++j) if (g == j % G) P.traj_out[D * (c0 + L) + j] = u[j];
        }
      }
      // ---------------- reverse sweep ------------------------------------------------------------------------------------
      if (GRAD) {
        double U0[D];
        int i = maxL - 1, n = S - 1;
        double t0 = 0.0, hs = 0.0;
        for (int k = nsteps - 1; k >= 0; --k) {
          const int b = k & 1;
          if (k > 0 && lane == 0) {      // prefetch step k-1 into the other buffer while step k is reversed
            mbar_expect_tx(&bar[b ^ 1], kStepBytes);
            bulk_load(s_stage + (b ^ 1) * kStepDoubles, stash + (long long)(k - 1) * rec_stride, kStepBytes, &bar[b ^ 1]);
          }
          mbar_wait(&bar[b], (bar_phase >> b) & 1u);
          bar_phase ^= (1u << b);
          double t;
          if (k >= pre) {
            if (n == S - 1) {
              const int ic = min(i, L);
              t0 = P.times[c0 + ic];
              hs = (i < L) ? (P.times[c0 + ic + 1] - t0) / S : 0.0;
            }
            t = fma((double)n, hs, t0);
          } else { t = P.times[c0] - eps; hs = is_null ? 0.0 : eps; }
          rk_reverse<R, G, HPL, SOLVER>(net, gr, ctx, lane, g, lam, t, hs, s_stage + b * kStepDoubles, U0, gseries);
          __syncwarp();                  // every lane is done reading buffer b before it is refilled
          if (k >= pre) {
            if (n == 0) {
              if (LK == LK_TRAJ) {   // U0 is the state at observation point i: add that point's loss cotangent
                const int ic = min(i, L);
                const double w = (i <= L && i >= first_scored) ? wpt : 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) lam[j] = fma(-2.0 * w, P.data[D * (c0 + ic) + j] - U0[j], lam[j]);
              }
              --i; n = S - 1;
            } else --n;
          }
        }
        if (!is_null) {
          if (LK == LK_STATE) {
            const long long c = sg.col;
#pragma unroll
            for (int j = 0; j < D; ++j) {
              if (g == j % G) {
                atomicAdd(guh + D * c + j, gc[j]);
                if (has_prev) atomicAdd(guh + D * c0 + j, gp[j] + lam[j]);
              }
            }
          } else if (from_theta) {
#pragma unroll
            for (int j = 0; j < D; ++j) if (g == j % G) atomicAdd(guh + D * c0 + j, lam[j]);
          }
        }
      }
    }
    if (GRAD && R::SERIES && !is_null && g == 0) {
      const long long ncm = P.model_col0[model + 1] - P.model_col0[model];
      atomicAdd(P.grad + P.theta_base[model] + (long long)D * ncm + P.off_series_param + series_local, gseries);
    }
  }
  // every block writes its partial row (zeros if it had no task) so the rows are always fully defined
  if (P.n_models == 1) flush_block<R, G, HPL, GRAD>(P, gr, loss, lane, g, s_red);
  else if (cur_model >= 0) flush_model<R, G, HPL, GRAD>(P, gr, loss, lane, g, cur_model);
}

}  // namespace ude
