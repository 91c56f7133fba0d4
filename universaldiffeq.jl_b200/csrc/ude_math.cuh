// FP64 scalar math used inside the UDE kernels (sm_100a).
//
// The MLP activation of every shipped constructor of the reference is tanh
// (/root/reference/src/ProcessModels.jl:379,427; src/NeuralNetworkConstructors.jl:20-22), which Lux
// lowers to NNlib.tanh_fast.  tanh dominates the FP64 issue budget of the hot
// path (one per hidden unit per RK stage), so it is hand-written: an expm1-style evaluation
//      tanh|x| = E / (E + 2),   E = expm1(2|x|) = 2^n (1 + p(r)) - 1,   2|x| = n ln2 + r
// with no cancellation anywhere (relative accuracy ~1 ulp-ish for all x, exact odd symmetry),
// a MUFU.RCP64H seeded Newton reciprocal instead of the IEEE division sequence, and no branches.
#pragma once
#include <cuda_runtime.h>

// Newton steps on the MUFU.RCP64H seed.  One step (2^-40) followed by the residual correction in ude_div / the tanh
// epilogue is already at the 2.2e-16 error floor (tools/microbench.cu built with -DUDE_RCP_NR=1, profiles/).
#ifndef UDE_RCP_NR
#define UDE_RCP_NR 1
#endif

// reciprocal of a normal double d (|d| in [2^-500, 2^500]); relative error ~1e-16 after the Newton steps
__device__ __forceinline__ double ude_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));      // MUFU.RCP64H: ~2^-20 relative
#pragma unroll
  for (int i = 0; i < UDE_RCP_NR; ++i) {
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
  }
  return y;
}

// num/den for den >= 1 (normal): one residual correction on top of ude_rcp
__device__ __forceinline__ double ude_div(double num, double den) {
  double y = ude_rcp(den);
  double q = num * y;
  double rem = fma(-q, den, num);
  return fma(rem, y, q);
}

__device__ __forceinline__ double ude_tanh(double x) {
  const double LOG2E  = 1.4426950408889634;
  const double LN2_HI = 6.93147180369123816490e-01;   // fdlibm split of ln 2
  const double LN2_LO = 1.90821492927058770002e-10;
  const double SHIFT  = 6755399441055744.0;           // 1.5 * 2^52: round-to-nearest-integer trick
  double t = fabs(x);
  t = t + t;
  t = fmin(t, 40.0);                                  // tanh(20) == 1.0 in double
  double kf = fma(t, LOG2E, SHIFT);
  int n = __double2loint(kf);
  kf -= SHIFT;
  double r = fma(kf, -LN2_HI, t);
  r = fma(kf, -LN2_LO, r);
  // p = expm1(r), |r| <= ln2/2: r + r^2 (1/2! + r/3! + ... + r^11/13!)
  double q = 1.6059043836821613e-10;                  // 1/13!
  q = fma(q, r, 2.08767569878681e-09);                // 1/12!
  q = fma(q, r, 2.505210838544172e-08);               // 1/11!
  q = fma(q, r, 2.755731922398589e-07);               // 1/10!
  q = fma(q, r, 2.7557319223985893e-06);              // 1/9!
  q = fma(q, r, 2.48015873015873e-05);                // 1/8!
  q = fma(q, r, 1.984126984126984e-04);               // 1/7!
  q = fma(q, r, 1.388888888888889e-03);               // 1/6!
  q = fma(q, r, 8.333333333333333e-03);               // 1/5!
  q = fma(q, r, 4.1666666666666664e-02);              // 1/4!
  q = fma(q, r, 1.6666666666666666e-01);              // 1/3!
  q = fma(q, r, 0.5);                                 // 1/2!
  double p = fma(q * r, r, r);
  double s = __hiloint2double((1023 + n) << 20, 0);   // 2^n, 0 <= n <= 58
  double em1 = fma(s, p, s - 1.0);                    // 2^n (1+p) - 1
  double y = ude_div(em1, em1 + 2.0);
  return copysign(y, x);
}

// 2^(j/64) table (512 B).  Kernels copy it to shared memory once per block; the lookup is one LDS on the
// otherwise idle LSU pipe and shortens the exp polynomial from degree 13 to 5, i.e. ~19 FP64-pipe
// instructions per tanh instead of ~33 (measured: tools/microbench.cu, profiles/microbench_r01.json).
__constant__ double c_ude_exp2_tab[64] = {
#include "ude_exp2_table.inc"
};

__device__ __forceinline__ void ude_load_exp2_tab(double* s_tab /* shared, 64 doubles */) {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_tab[i] = c_ude_exp2_tab[i];
}
// COPIES interleaved copies: entry j of copy q at [j * COPIES + q]
template <int COPIES>
__device__ __forceinline__ void ude_load_exp2_tab_n(double* s_tab /* shared, 64 * COPIES doubles */) {
  for (int i = threadIdx.x; i < 64 * COPIES; i += blockDim.x) s_tab[i] = c_ude_exp2_tab[i / COPIES];
}

// tanh via E = expm1(2|x|) = 2^n T_j (1 + p(r)) - 1 with 2|x| = (64 n + j) ln2/64 + r, |r| <= ln2/128.
// |error| <= ~1.5e-16 absolute everywhere (relative ~1e-16 except for |x| in [0.003, 0.3] where E = T(1+p)-1
// cancels mildly: still <= 2^-52 absolute).  Branch-free; the clamp to |x| <= 20 is done on the high word
// in the integer pipe.
__device__ __forceinline__ double ude_tanh_tab(double x, const double* __restrict__ s_tab) {
  const double INV   = 184.6649652337873;              // 128 / ln 2
  const double C_HI  = 0x1.62e42fee00000p-7;           // (ln 2)/64, high part (21 trailing zero bits)
  const double C_LO  = 0x1.a39ef35793c76p-39;
  const double SHIFT = 6755399441055744.0;
  int hi = __double2hiint(x) & 0x7fffffff;
  hi = min(hi, 0x40340000);                            // |x| <= 20 (+ < 2e-6 when the low word is kept)
  double ax = __hiloint2double(hi, __double2loint(x));
  double kf = fma(ax, INV, SHIFT);
  int N = __double2loint(kf);
  kf -= SHIFT;
  double r = fma(kf, -C_HI, ax + ax);
  r = fma(kf, -C_LO, r);
  double q = fma(r, 8.333333333333333e-03, 4.1666666666666664e-02);
  q = fma(q, r, 1.6666666666666666e-01);
  q = fma(q, r, 0.5);
  double p = fma(q, r * r, r);                         // expm1(r)
  double T = s_tab[N & 63];
  double s = __hiloint2double(__double2hiint(T) + ((N >> 6) << 20), __double2loint(T));   // 2^n T_j
  double em1 = fma(s, p, s - 1.0);
  double y = ude_div(em1, em1 + 2.0);
  return copysign(y, x);
}

// The same evaluation for N independent arguments, written phase by phase so that the N dependency chains are
// interleaved in program order (the FP64 pipe has an 8-cycle dependent-issue latency and issues every 2 cycles:
// a single chain leaves it 75% idle; profiles/ncu_c3_r01a.txt shows stall_wait as the top stall reason).
template <int N>
__device__ __forceinline__ void ude_tanh_tab_vec(const double (&x)[N], double (&out)[N], const double* __restrict__ s_tab) {
  const double INV   = 184.6649652337873;
  const double C_HI  = 0x1.62e42fee00000p-7;
  const double C_LO  = 0x1.a39ef35793c76p-39;
  const double SHIFT = 6755399441055744.0;
  double ax[N], kf[N], r[N], q[N], p[N], T[N], s[N], em1[N], den[N], y[N], e[N];
  int Nn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    int hi = __double2hiint(x[i]) & 0x7fffffff;
    hi = min(hi, 0x40340000);
    ax[i] = __hiloint2double(hi, __double2loint(x[i]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) kf[i] = fma(ax[i], INV, SHIFT);
#pragma unroll
  for (int i = 0; i < N; ++i) { Nn[i] = __double2loint(kf[i]); kf[i] -= SHIFT; }
#pragma unroll
  for (int i = 0; i < N; ++i) T[i] = s_tab[Nn[i] & 63];
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(kf[i], -C_HI, ax[i] + ax[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(kf[i], -C_LO, r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(r[i], 8.333333333333333e-03, 4.1666666666666664e-02);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], r[i], 1.6666666666666666e-01);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], r[i], 0.5);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(q[i], r[i] * r[i], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = __hiloint2double(__double2hiint(T[i]) + ((Nn[i] >> 6) << 20), __double2loint(T[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) em1[i] = fma(s[i], p[i], s[i] - 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) den[i] = em1[i] + 2.0;
#pragma unroll
  for (int i = 0; i < N; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(den[i]));
#pragma unroll
  for (int it = 0; it < UDE_RCP_NR; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = fma(-den[i], y[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = em1[i] * y[i];
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-q[i], den[i], em1[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = copysign(fma(e[i], y[i], q[i]), x[i]);
}

// Round-2 evaluation: 14 FP64-pipe instructions per tanh instead of 19 (tanh is ~30% of the FP64 issue of the flagship
// kernel).  With w = exp(2|x|) = 2^n T_j (1 + p):   tanh|x| = 1 - 2 / (w + 1).
//   * argument reduction on |x| directly (no doubling) and with ONE constant: rh = |x| - k (ln2/128) = r/2.  The rounding
//     of the constant puts an error of 1.1e-16 |x| on r, which the result sees scaled by 2w/(w+1)^2 <= e^(-2|x|)/2...1/2:
//     below 3e-17 for every x, so the second Cody-Waite term of ude_tanh_tab buys nothing here;
//   * P = p/2 = rh + rh^2 (1 + 2/3 rh + 1/3 rh^2 + 2/15 rh^3)  (expm1(2 rh)/2 to r^5; the r^6/720 term is 3.4e-17);
//   * w + 1 = fma(2s, P, s + 1) with s = 2^n T_j assembled in the integer pipe;
//   * reciprocal by MUFU.RCP64H (~2^-20) and ONE cubically convergent step y(1 + e + e^2): (2^-20)^3 = 2^-60;
//   * 1 - 2y with a single fma.  Absolute error <= 2.2e-16 (same floor as ude_tanh_tab; relative accuracy is lost only
//     where |tanh| < 1e-6, where the absolute floor is what the sums this feeds can see anyway).
// TS / toff: table stride and offset -- TS = 16, toff = lane & 15 reads an interleaved 16-copy table without bank conflicts
template <int N, int TS = 1>
__device__ __forceinline__ void ude_tanh_fast_vec(const double (&x)[N], double (&out)[N], const double* __restrict__ s_tab, int toff = 0) {
  const double INV   = 184.6649652337873;              // 128 / ln 2
  const double CH    = 0x1.62e42fefa39efp-8;           // (ln 2)/128
  const double SHIFT = 6755399441055744.0;
  double ax[N], kf[N], rh[N], q[N], P[N], T[N], s[N], s2[N], den[N], y[N], e[N];
  int Nn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    int hi = __double2hiint(x[i]) & 0x7fffffff;
    hi = min(hi, 0x40340000);                          // |x| <= 20
    ax[i] = __hiloint2double(hi, __double2loint(x[i]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) kf[i] = fma(ax[i], INV, SHIFT);
#pragma unroll
  for (int i = 0; i < N; ++i) { Nn[i] = __double2loint(kf[i]); kf[i] -= SHIFT; }
#pragma unroll
  for (int i = 0; i < N; ++i) T[i] = s_tab[(Nn[i] & 63) * TS + toff];
#pragma unroll
  for (int i = 0; i < N; ++i) rh[i] = fma(kf[i], -CH, ax[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(rh[i], 1.3333333333333333e-01, 3.3333333333333331e-01);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 6.6666666666666663e-01);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) P[i] = fma(q[i], rh[i] * rh[i], rh[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int hi = __double2hiint(T[i]) + ((Nn[i] >> 6) << 20);
    s[i] = __hiloint2double(hi, __double2loint(T[i]));                     // 2^n T_j
    s2[i] = __hiloint2double(hi + (1 << 20), __double2loint(T[i]));        // 2 s
  }
#pragma unroll
  for (int i = 0; i < N; ++i) den[i] = fma(s2[i], P[i], s[i] + 1.0);       // w + 1
#pragma unroll
  for (int i = 0; i < N; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(den[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-den[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = copysign(fma(-2.0, y[i], 1.0), x[i]);
}

// Variant with 13 FP64-pipe instructions: since the result is formed as 1 - 2/(w + 1), w = exp(2|x|) needs no expm1-style
// care, so the polynomial is a plain Horner form of E = exp(2 rh) (degree 5, |2 rh| <= ln2/128: the r^6/720 term is 3.4e-17)
// and  w + 1 = fma(s, E, 1)  -- one FMA fewer than P / s2 / s + 1, and no second exponent patch in the integer pipe.
// The rounding of E and of the fma put <= 2.5e-16 relative on w + 1, i.e. <= 2.5e-16 * 2w/(w+1)^2 <= 1.3e-16 absolute on
// the result, on top of the same reciprocal / final-fma roundings as ude_tanh_fast_vec (tools/tanh_check.cu measures it).
template <int N, int TS = 1>
__device__ __forceinline__ void ude_tanh_fast2_vec(const double (&x)[N], double (&out)[N], const double* __restrict__ s_tab, int toff = 0) {
  const double INV   = 184.6649652337873;              // 128 / ln 2
  const double CH    = 0x1.62e42fefa39efp-8;           // (ln 2)/128
  const double SHIFT = 6755399441055744.0;
  double ax[N], kf[N], rh[N], q[N], T[N], s[N], den[N], y[N], e[N];
  int Nn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    int hi = __double2hiint(x[i]) & 0x7fffffff;
    hi = min(hi, 0x40340000);                          // |x| <= 20
    ax[i] = __hiloint2double(hi, __double2loint(x[i]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) kf[i] = fma(ax[i], INV, SHIFT);
#pragma unroll
  for (int i = 0; i < N; ++i) { Nn[i] = __double2loint(kf[i]); kf[i] -= SHIFT; }
#pragma unroll
  for (int i = 0; i < N; ++i) T[i] = s_tab[(Nn[i] & 63) * TS + toff];
#pragma unroll
  for (int i = 0; i < N; ++i) rh[i] = fma(kf[i], -CH, ax[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(rh[i], 2.6666666666666666e-01, 6.6666666666666663e-01);   // 4/15, 2/3
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 1.3333333333333333);                          // 4/3
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 2.0);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 2.0);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 1.0);                                         // exp(2 rh)
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = __hiloint2double(__double2hiint(T[i]) + ((Nn[i] >> 6) << 20), __double2loint(T[i]));   // 2^n T_j
#pragma unroll
  for (int i = 0; i < N; ++i) den[i] = fma(s[i], q[i], 1.0);                                        // w + 1
#pragma unroll
  for (int i = 0; i < N; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(den[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-den[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = copysign(fma(-2.0, y[i], 1.0), x[i]);
}

// Variant 3: the same 13 FP64-pipe instructions with ~4 fewer integer / move instructions around them (the tanh phases
// of the kernels are ISSUE-bound, DESIGN.md 4.0, so every instruction counts, not only the FP64 ones):
//   * SIGNED argument: w = exp(2x), tanh x = 1 - 2/(w + 1) holds for either sign, so |x| and copysign go;
//   * the clamp |x| <= 20 (beyond which the result is exactly +-1) is two integer min on the high word, in place: a signed
//     min against hi(20) bounds the positive arguments (negative ones are negative integers), an unsigned min against
//     hi(-20) bounds the negative ones (positive ones have the top bit clear).  +-inf saturate to +-1; a NaN becomes 1
//     (as in the other variants: a diverged state shows up as a non-finite loss through the state itself);
//   * BIASED table (ude_load_exp2_tabb*): entry j carries hi(T_j) - (j << 14), so the exponent patch is the single
//     integer operation hi + (N << 14) = hi(T_j) + (n << 20), N = 64 n + j (no masking of j, no second shift).
// Not exactly odd (tanh(-x) and -tanh(x) may differ in the last bit); absolute error as variant 2 (tools/tanh_check.cu).
template <int N, int TS = 1>
__device__ __forceinline__ void ude_tanh_fast3_vec(const double (&x)[N], double (&out)[N], const double* __restrict__ s_tabb, int toff = 0) {
  const double INV   = 184.6649652337873;              // 128 / ln 2
  const double CH    = 0x1.62e42fefa39efp-8;           // (ln 2)/128
  const double SHIFT = 6755399441055744.0;
  double xc[N], kf[N], rh[N], q[N], T[N], s[N], den[N], y[N], e[N];
  int Nn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    int hi = min(__double2hiint(x[i]), 0x40340000);
    hi = (int)min((unsigned)hi, 0xc0340000u);
    xc[i] = __hiloint2double(hi, __double2loint(x[i]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) kf[i] = fma(xc[i], INV, SHIFT);
#pragma unroll
  for (int i = 0; i < N; ++i) { Nn[i] = __double2loint(kf[i]); kf[i] -= SHIFT; }
#pragma unroll
  for (int i = 0; i < N; ++i) T[i] = s_tabb[(Nn[i] & 63) * TS + toff];
#pragma unroll
  for (int i = 0; i < N; ++i) rh[i] = fma(kf[i], -CH, xc[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(rh[i], 2.6666666666666666e-01, 6.6666666666666663e-01);   // 4/15, 2/3
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 1.3333333333333333);                          // 4/3
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 2.0);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 2.0);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 1.0);                                         // exp(2 rh)
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = __hiloint2double(__double2hiint(T[i]) + (int)((unsigned)Nn[i] << 14), __double2loint(T[i]));   // 2^n T_j
#pragma unroll
  for (int i = 0; i < N; ++i) den[i] = fma(s[i], q[i], 1.0);                                        // w + 1
#pragma unroll
  for (int i = 0; i < N; ++i) {
    // MUFU.RCP64H produces the high word only; the seed (~2^-20) does not care about the low one, so it borrows the low word
    // of s (dead from here on) instead of paying a move for a zero
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(den[i]));
    y[i] = __hiloint2double(__double2hiint(y0), __double2loint(s[i]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-den[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = fma(-2.0, y[i], 1.0);
}

// biased 2^(j/64) table for ude_tanh_fast3_vec: COPIES interleaved copies, entry j of copy q at [j * COPIES + q]
template <int COPIES>
__device__ __forceinline__ void ude_load_exp2_tabb_n(double* s_tab /* shared, 64 * COPIES doubles */) {
  for (int i = threadIdx.x; i < 64 * COPIES; i += blockDim.x) {
    const int j = i / COPIES;
    const double T = c_ude_exp2_tab[j];
    s_tab[i] = __hiloint2double(__double2hiint(T) - (j << 14), __double2loint(T));
  }
}

// Variant 4 (kernels whose table is a single copy): a 256-entry table, |rh| <= ln2/1024, so exp(2 rh) needs degree 4 only
// (the (2 rh)^5/120 term is 3.8e-17): 12 FP64-pipe instructions and a dependent chain one FMA shorter.  Same structure as
// variant 3: k = round(x 512/ln2) = 256 n + j, biased entry hi(T_j) - (j << 12), exponent patch hi + (N << 12).  The single
// reduction constant ln2/512 has the same relative rounding as ln2/128, so the damping argument of ude_tanh_fast_vec holds.
__constant__ double c_ude_exp2_tab256[256] = {
#include "ude_exp2_table256.inc"
};
__device__ __forceinline__ void ude_load_exp2_tabb256(double* s_tab /* shared, 256 doubles */) {
  for (int j = threadIdx.x; j < 256; j += blockDim.x) {
    const double T = c_ude_exp2_tab256[j];
    s_tab[j] = __hiloint2double(__double2hiint(T) - (j << 12), __double2loint(T));
  }
}
template <int N>
__device__ __forceinline__ void ude_tanh_fast4_vec(const double (&x)[N], double (&out)[N], const double* __restrict__ s_tabb) {
  const double INV   = 738.6598609351493;               // 512 / ln 2
  const double CH    = 0x1.62e42fefa39efp-10;           // (ln 2)/512
  const double SHIFT = 6755399441055744.0;
  double xc[N], kf[N], rh[N], q[N], T[N], s[N], den[N], y[N], e[N];
  int Nn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    int hi = min(__double2hiint(x[i]), 0x40340000);
    hi = (int)min((unsigned)hi, 0xc0340000u);
    xc[i] = __hiloint2double(hi, __double2loint(x[i]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) kf[i] = fma(xc[i], INV, SHIFT);
#pragma unroll
  for (int i = 0; i < N; ++i) { Nn[i] = __double2loint(kf[i]); kf[i] -= SHIFT; }
#pragma unroll
  for (int i = 0; i < N; ++i) T[i] = s_tabb[Nn[i] & 255];
#pragma unroll
  for (int i = 0; i < N; ++i) rh[i] = fma(kf[i], -CH, xc[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(rh[i], 6.6666666666666663e-01, 1.3333333333333333);       // 2/3, 4/3
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 2.0);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 2.0);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], rh[i], 1.0);                                         // exp(2 rh)
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = __hiloint2double(__double2hiint(T[i]) + (int)((unsigned)Nn[i] << 12), __double2loint(T[i]));   // 2^n T_j
#pragma unroll
  for (int i = 0; i < N; ++i) den[i] = fma(s[i], q[i], 1.0);                                        // w + 1
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(den[i]));
    y[i] = __hiloint2double(__double2hiint(y0), __double2loint(s[i]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-den[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = fma(-2.0, y[i], 1.0);
}

// type-dispatched front end used by the kernels (lane-per-hidden-unit family and the round-1 DMMA family; the round-2 DMMA
// kernel picks its variant by option flag).  UDE_TANH_V: 4 = ude_tanh_fast4_vec (default), 3 = fast3, 2 = fast2, 1 = fast,
// 0 = tab.  The table a kernel stages in shared memory must match: UDE_TANH_TAB_N doubles filled by ude_load_tanh_tab.
#ifndef UDE_TANH_V
#define UDE_TANH_V 4
#endif
#define UDE_TANH_TAB_N (UDE_TANH_V == 4 ? 256 : 64)
__device__ __forceinline__ void ude_load_tanh_tab(double* s_tab /* shared, UDE_TANH_TAB_N doubles */) {
#if UDE_TANH_V == 4
  ude_load_exp2_tabb256(s_tab);
#elif UDE_TANH_V == 3
  ude_load_exp2_tabb_n<1>(s_tab);
#else
  ude_load_exp2_tab(s_tab);
#endif
}
template <int N>
__device__ __forceinline__ void ude_tanh_vec(const double (&x)[N], double (&out)[N], const double* __restrict__ s_tab) {
#if UDE_TANH_V == 4
  ude_tanh_fast4_vec<N>(x, out, s_tab);
#elif UDE_TANH_V == 3
  ude_tanh_fast3_vec<N>(x, out, s_tab);
#elif UDE_TANH_V == 2
  ude_tanh_fast2_vec<N>(x, out, s_tab);
#elif UDE_TANH_V == 1
  ude_tanh_fast_vec<N>(x, out, s_tab);
#else
  ude_tanh_tab_vec<N>(x, out, s_tab);
#endif
}

// single-precision twin for the optional FP32 mode (1e-4 parity bar): hardware tanh is too coarse
// (2^-11), so use the same E/(E+2) form with ex2.approx; ~1e-6 relative.
__device__ __forceinline__ float ude_tanhf(float x) {
  float t = fminf(fabsf(x) * 2.0f, 20.0f);
  float e = __expf(t);
  float y = __fdividef(e - 1.0f, e + 1.0f);
  // small-|x| branch-free fix: for t < 2^-5 use the odd Taylor polynomial (avoids e-1 cancellation)
  float x2 = x * x;
  float ys = fabsf(x) * (1.0f + x2 * (-0.33333334f + x2 * 0.13333334f));
  y = (t < 0.0625f) ? ys : y;
  return copysignf(y, x);
}

template <int N>
__device__ __forceinline__ void ude_tanh_vec(const float (&x)[N], float (&out)[N], const double* __restrict__) {
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = ude_tanhf(x[i]);
}
