// Right-hand-side catalog (device side).  Each trait states the sizes the kernel is specialised on and
// three pure functions: how the MLP input is assembled, how the MLP output is combined with the
// parametric terms into du, and the vector-Jacobian product of that combination.  The shapes are the ones
// the reference's constructors and docs produce:
//   RhsNode<D,NX>            NODE / MultiNODE (/root/reference/src/ProcessModels.jl:387-390, :434-437;
//                            src/MultiProcessModel.jl:215-218, :255-258)
//   RhsNodeTime<D,NX,DOUT>   time-dependent NODE via CustomDerivatives (docs/src/examples.md:44-49)
//   RhsLvUde<NX>             test/UDETests.jl:20-28 (NX=0) and :82-86 (NX=1, + beta[i]*X[1])
//   RhsLvUdeAbs              README.md:52-62
//   RhsFishery               docs/src/examples.md:163-166
//   RhsNnDecay<D>            test/MultiUDE.jl:17-19
//   RhsSeriesGrowth<D>       docs/src/MultipleTimeSeries.md:51-55
// All values here are uniform across the lanes of a segment's group (every lane computes them redundantly).
#pragma once
#include "../../include/ude_cuda.h"

namespace ude {

struct RhsCtx {
  double sc[UDE_MAX_SCALARS];   // scalar parameters of the model (uniform registers)
  double series_param;          // r[series] for RhsSeriesGrowth
  double time_scale, time_shift;
};

__device__ __forceinline__ double sgn1(double v) { return v < 0.0 ? -1.0 : 1.0; }

template <int D_, int NX_>
struct RhsNode {
  static constexpr int KIND = UDE_RHS_NODE, D = D_, NX = NX_, DIN = D_ + NX_, DOUT = D_, NS = 0;
  static constexpr bool NEED_Y = false, SERIES = false;
  __device__ __forceinline__ static void input(const double (&U)[D], const double* X, double, const RhsCtx&, double (&x)[DIN]) {
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = U[j];
#pragma unroll
    for (int v = 0; v < NX; ++v) x[D + v] = X[v];
  }
  __device__ __forceinline__ static void combine(const double (&)[D], const double*, const double (&y)[DOUT], const RhsCtx&, double (&du)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) du[j] = y[j];
  }
  __device__ __forceinline__ static void combine_vjp(const double (&)[D], const double*, const double (&)[DOUT], const RhsCtx&,
                                                     const double (&kb)[D], double (&yb)[DOUT], double (&)[D], double*, double&) {
#pragma unroll
    for (int j = 0; j < D; ++j) yb[j] = kb[j];
  }
};

template <int D_, int NX_, int DOUT_>
struct RhsNodeTime {
  static constexpr int KIND = UDE_RHS_NODE_TIME, D = D_, NX = NX_, DIN = D_ + NX_ + 1, DOUT = DOUT_, NS = D_ - DOUT_;
  static constexpr bool NEED_Y = false, SERIES = false;
  __device__ __forceinline__ static void input(const double (&U)[D], const double* X, double t, const RhsCtx& c, double (&x)[DIN]) {
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = U[j];
#pragma unroll
    for (int v = 0; v < NX; ++v) x[D + v] = X[v];
    x[D + NX] = c.time_scale * t + c.time_shift;
  }
  __device__ __forceinline__ static void combine(const double (&U)[D], const double*, const double (&y)[DOUT], const RhsCtx& c, double (&du)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) du[j] = (j < DOUT) ? y[j < DOUT ? j : 0] : -c.sc[j >= DOUT ? j - DOUT : 0] * U[j];
  }
  __device__ __forceinline__ static void combine_vjp(const double (&U)[D], const double*, const double (&)[DOUT], const RhsCtx& c,
                                                     const double (&kb)[D], double (&yb)[DOUT], double (&ub)[D], double* gsc, double&) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (j < DOUT) yb[j < DOUT ? j : 0] = kb[j];
      else { ub[j] += -c.sc[j >= DOUT ? j - DOUT : 0] * kb[j]; gsc[j >= DOUT ? j - DOUT : 0] += -U[j] * kb[j]; }
    }
  }
};

template <int NX_>
struct RhsLvUde {   // scalars: r, m, theta (, beta1, beta2)
  static constexpr int KIND = UDE_RHS_LV_UDE, D = 2, NX = NX_, DIN = 2, DOUT = 1, NS = NX_ ? 5 : 3;
  static constexpr bool NEED_Y = true, SERIES = false;
  __device__ __forceinline__ static void input(const double (&U)[D], const double*, double, const RhsCtx&, double (&x)[DIN]) {
    x[0] = U[0]; x[1] = U[1];
  }
  __device__ __forceinline__ static void combine(const double (&U)[D], const double* X, const double (&y)[DOUT], const RhsCtx& c, double (&du)[D]) {
    du[0] = c.sc[0] * U[0] - y[0];
    du[1] = c.sc[2] * y[0] - c.sc[1] * U[1];
    if (NX) { du[0] += c.sc[3] * X[0]; du[1] += c.sc[4] * X[0]; }
  }
  __device__ __forceinline__ static void combine_vjp(const double (&U)[D], const double* X, const double (&y)[DOUT], const RhsCtx& c,
                                                     const double (&kb)[D], double (&yb)[DOUT], double (&ub)[D], double* gsc, double&) {
    yb[0] = -kb[0] + c.sc[2] * kb[1];
    ub[0] += c.sc[0] * kb[0];
    ub[1] += -c.sc[1] * kb[1];
    gsc[0] += U[0] * kb[0]; gsc[1] += -U[1] * kb[1]; gsc[2] += y[0] * kb[1];
    if (NX) { gsc[3] += X[0] * kb[0]; gsc[4] += X[0] * kb[1]; }
  }
};

struct RhsLvUdeAbs {   // scalars (NamedTuple order of README.md:49): r, k, m, theta
  static constexpr int KIND = UDE_RHS_LV_UDE_ABS, D = 2, NX = 0, DIN = 2, DOUT = 1, NS = 4;
  static constexpr bool NEED_Y = true, SERIES = false;
  __device__ __forceinline__ static void input(const double (&U)[D], const double*, double, const RhsCtx&, double (&x)[DIN]) {
    x[0] = U[0]; x[1] = U[1];
  }
  __device__ __forceinline__ static void combine(const double (&U)[D], const double*, const double (&y)[DOUT], const RhsCtx& c, double (&du)[D]) {
    double r = fabs(c.sc[0]), k = fabs(c.sc[1]), m = fabs(c.sc[2]), th = fabs(c.sc[3]), C = fabs(y[0]);
    du[0] = r * U[0] * (1.0 - U[0] / k) - C;
    du[1] = th * C - m * U[1];
  }
  __device__ __forceinline__ static void combine_vjp(const double (&U)[D], const double*, const double (&y)[DOUT], const RhsCtx& c,
                                                     const double (&kb)[D], double (&yb)[DOUT], double (&ub)[D], double* gsc, double&) {
    double r = fabs(c.sc[0]), k = fabs(c.sc[1]), m = fabs(c.sc[2]), th = fabs(c.sc[3]), C = fabs(y[0]);
    yb[0] = sgn1(y[0]) * (-kb[0] + th * kb[1]);
    ub[0] += r * (1.0 - 2.0 * U[0] / k) * kb[0];
    ub[1] += -m * kb[1];
    gsc[0] += sgn1(c.sc[0]) * U[0] * (1.0 - U[0] / k) * kb[0];
    gsc[1] += sgn1(c.sc[1]) * r * U[0] * U[0] / (k * k) * kb[0];
    gsc[2] += sgn1(c.sc[2]) * (-U[1]) * kb[1];
    gsc[3] += sgn1(c.sc[3]) * C * kb[1];
  }
};

struct RhsFishery {   // scalars: q, r, K ; NN([u; X]) 3 -> H -> 1
  static constexpr int KIND = UDE_RHS_FISHERY, D = 2, NX = 1, DIN = 3, DOUT = 1, NS = 3;
  static constexpr bool NEED_Y = false, SERIES = false;
  __device__ __forceinline__ static void input(const double (&U)[D], const double* X, double, const RhsCtx&, double (&x)[DIN]) {
    x[0] = U[0]; x[1] = U[1]; x[2] = X[0];
  }
  __device__ __forceinline__ static void combine(const double (&U)[D], const double*, const double (&y)[DOUT], const RhsCtx& c, double (&du)[D]) {
    du[0] = y[0];
    du[1] = c.sc[1] * U[1] * (1.0 - U[1] / c.sc[2]) - c.sc[0] * U[0];
  }
  __device__ __forceinline__ static void combine_vjp(const double (&U)[D], const double*, const double (&)[DOUT], const RhsCtx& c,
                                                     const double (&kb)[D], double (&yb)[DOUT], double (&ub)[D], double* gsc, double&) {
    double q = c.sc[0], r = c.sc[1], K = c.sc[2];
    yb[0] = kb[0];
    ub[0] += -q * kb[1];
    ub[1] += r * (1.0 - 2.0 * U[1] / K) * kb[1];
    gsc[0] += -U[0] * kb[1];
    gsc[1] += U[1] * (1.0 - U[1] / K) * kb[1];
    gsc[2] += r * U[1] * U[1] / (K * K) * kb[1];
  }
};

template <int D_>
struct RhsNnDecay {   // scalars: m
  static constexpr int KIND = UDE_RHS_NN_DECAY, D = D_, NX = 0, DIN = D_, DOUT = D_, NS = 1;
  static constexpr bool NEED_Y = false, SERIES = false;
  __device__ __forceinline__ static void input(const double (&U)[D], const double*, double, const RhsCtx&, double (&x)[DIN]) {
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = U[j];
  }
  __device__ __forceinline__ static void combine(const double (&U)[D], const double*, const double (&y)[DOUT], const RhsCtx& c, double (&du)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) du[j] = y[j] - c.sc[0] * U[j];
  }
  __device__ __forceinline__ static void combine_vjp(const double (&U)[D], const double*, const double (&)[DOUT], const RhsCtx& c,
                                                     const double (&kb)[D], double (&yb)[DOUT], double (&ub)[D], double* gsc, double&) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) { yb[j] = kb[j]; ub[j] += -c.sc[0] * kb[j]; acc += U[j] * kb[j]; }
    gsc[0] += -acc;
  }
};

template <int D_>
struct RhsSeriesGrowth {   // per-series r[i]; NN(u) D -> H -> 1
  static constexpr int KIND = UDE_RHS_SERIES_GROWTH, D = D_, NX = 0, DIN = D_, DOUT = 1, NS = 0;
  static constexpr bool NEED_Y = true, SERIES = true;
  __device__ __forceinline__ static void input(const double (&U)[D], const double*, double, const RhsCtx&, double (&x)[DIN]) {
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = U[j];
  }
  __device__ __forceinline__ static void combine(const double (&U)[D], const double*, const double (&y)[DOUT], const RhsCtx& c, double (&du)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) du[j] = c.series_param * U[j] * y[0];
  }
  __device__ __forceinline__ static void combine_vjp(const double (&U)[D], const double*, const double (&y)[DOUT], const RhsCtx& c,
                                                     const double (&kb)[D], double (&yb)[DOUT], double (&ub)[D], double*, double& gseries) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) { acc += U[j] * kb[j]; ub[j] += c.series_param * y[0] * kb[j]; }
    yb[0] = c.series_param * acc;
    gseries += y[0] * acc;
  }
};

}  // namespace ude
