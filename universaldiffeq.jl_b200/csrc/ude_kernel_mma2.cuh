// Round-2 tensor-core kernel for NODE-type right-hand sides whose hidden layer fits one warp (H <= 8*HT): the same
// fragment algebra as ude_kernel_mma.cuh (mma.sync.m8n8k4.f64 / DMMA, eight segments per M tile, no shuffle in the stage
// loop, weight gradients as MMAs over the batch, TMA-staged stash of the stage records) with the three changes the
// round-1 profile asked for (profiles/ncu_c3_mma_r01g.txt, VERDICT round 1 "next" item 1):
//   (a) NO dependent global load inside the step loops.  The times, data columns and covariate knots an observation
//       interval needs are loaded into registers ONE INTERVAL AHEAD (the loads are issued when the previous interval
//       starts, ~16 RK stages before their first use), and covariates are advanced knot by knot from those prefetched
//       values -- the per-crossing binary search over global knots (5 dependent loads, the long_sb DSETP/DADD hot spots)
//       is gone; one search per segment remains at task start.  Covariate evaluation is one DADD + one DFMA
//       (x = x_hi + slope (t - t_hi)), and stages that share a time (RK4 c2 = c3) share the value.
//   (b) NT = 2 M-tiles (16 segments) per warp: every weight B fragment read from shared memory (LDS.64) feeds two DMMAs,
//       and a warp carries twice the independent DMMA / tanh chains, so two resident warps per scheduler keep the FP64
//       pipe fed where three were not.
//   (c) fewer FP64-pipe instructions around the MMAs: tanh in 14 instead of 19 (ude_tanh_fast_vec; MMA2_TANH3: 13 + 8 other), b2 rides in the
//       accumulator of the first layer-2 DMMA, step sizes are formed with a multiply by 1/substeps.
// Loss families LK_STATE and LK_TRAJ (derivative matching has no time stepping: ude_kernel_mma.cuh keeps it).
// Reference semantics: src/MultiProcessModel.jl:215-218 (RHS), src/covariates.jl:46-48 (interpolation),
// LFC.jl:261-305, :639-708 (multiple shooting), :196-232, :545-601 (shooting), :17-98, :347-404 (conditional likelihood).
#pragma once
#include "ude_kernel_mma.cuh"

namespace ude {

// OPT bits (compile-time tuning knobs, measured in profiles/sweep_r02*.json)
enum { MMA2_PARK_GW = 1,      // park the weight-gradient accumulators in shared memory during the forward sweep (they are
                              // only touched by the reverse sweep): ~36 fewer live registers where pressure peaks (tanh)
       MMA2_TAB16 = 2,
       MMA2_DIRECT_ST = 4,    // forward sweep writes the stage records straight from registers to the stash (STG.64, whole
                              // 32-byte sectors) instead of staging them in shared memory for a TMA bulk store: no
                              // fence.proxy.async / __syncwarp / wait_group per stage; the reverse sweep still bulk-loads
       MMA2_NO_PIN = 8,       // A/B: plain loads for the interval prefetch (ptxas sinks them to the use)
       MMA2_SMEM_PF = 32,     // the one-interval-ahead prefetch of times / data / covariate knots goes global -> shared memory with
                              // cp.async (LDGSTS) and is read back with LDS when needed, instead of living in registers: register
                              // prefetches share the 6 scoreboards of a warp with every other pending load, so their consumers
                              // still showed long_sb stalls (profiles/ncu_c3_mma2_r02b.txt, DSETP at the covariate check)
       MMA2_AFFINE = 64,      // D a multiple of 4: the covariate and constant-1 input slots leave the layer-1 MMA (they would cost a
                              // whole extra k-step for 1 + NX useful slots of 4); the accumulator starts at b1 + sum_q x_q W1[:, D+q]
                              // instead (NX DFMAs per hidden unit, weights read as LDS.128 pairs): 4 DMMAs -> 8 DFMAs per stage
       MMA2_ROT = 16,         // A/B: round-1 [segment][hidden] tiles with the column rotated by 4*segment (dense rows, 11% less
                              // stash traffic) instead of rows padded by 4 doubles.  Padded rows make every shared-memory
                              // address of the stage loop "lane base + immediate": the rotation costs two integer
                              // instructions per access (~60 of the ~590 issued per stage and tile)
       // ablation bits for timing diagnosis ONLY (results are wrong): registered by ude_inst_mma2v_probe.cu
       MMA2_TANH2 = 8192,     // 13-instruction tanh (ude_tanh_fast2_vec)
       MMA2_RING3 = 32768,    // reverse sweep: three-slot record ring, the record TWO stages ahead is in flight (one stage of lead is about
                              // one HBM round trip with three warps per scheduler: the mbarrier waits showed up as long_sb stalls)
       MMA2_TANH3 = 16384,    // + signed argument, integer clamp, biased table (ude_tanh_fast3_vec): ~4 fewer non-FP64 instructions
       MMA2_ABL_NO_TMA = 256, MMA2_ABL_NO_TANH = 512, MMA2_ABL_NO_GW = 1024, MMA2_ABL_NO_REC = 2048, MMA2_ABL_NO_COV = 4096 };      // 16 interleaved copies of the 2^(j/64) table: lane l of a half-warp reads bank pair l, so the
                              // table lookups of tanh are conflict-free (8 KB per block instead of 512 B)

template <int D_, int NX_, int HT_, int NT_, int OPT_ = 0, int TB_ = 4>
struct Mma2Geo : MmaGeo<D_, NX_, HT_, 1, 4> {
  using Base = MmaGeo<D_, NX_, HT_, 1, 4>;
  static constexpr int NT = NT_;                       // M tiles (8 segments each) per warp
  static constexpr int OPT = OPT_, TB = TB_;
  static constexpr int KSU = (D_ + 3) / 4;             // slot steps that can hold state components
  // [slot][segment] x tile and [output][segment] k-bar tile: the owner's stores (lane (r, c) -> slot c + 4e, segment r)
  // and the transposed A-fragment loads of the weight-gradient MMAs (slot r, segment 4 s2 + c) both touch 32 distinct
  // bank pairs (the [segment][slot] layout of ude_mma_kernel had 4-way / 2-way conflicts, profiles/ncu_c3_mma2nt1_r02a.txt)
  __host__ __device__ static constexpr int xpos(int seg, int slot) { return slot * 8 + seg; }
  __host__ __device__ static constexpr int kpos(int seg, int out) { return out * 8 + seg; }
  static constexpr bool ROT = (OPT_ & 16) != 0 && Base::HPOW2;
  static constexpr int HS = ROT ? 8 * HT_ : 8 * HT_ + 4;     // row stride of the [segment][hidden] tiles (hides Base::HS)
  __host__ __device__ static constexpr int hpos(int seg, int col) {
    return ROT ? seg * HS + ((col + 4 * seg) & (HS - 1)) : seg * HS + col;
  }
  static constexpr int REC1 = 8 * HS + 8 * Base::XS;   // doubles per stage record of one tile (h tile + x tile)
  static constexpr int RECW = NT_ * REC1;              // per warp and stage
  static constexpr int SCR1 = 8 * HS + 64;             // z-bar tile + k-bar tile of one M tile
  static constexpr int NGW = (Base::MT + 1) * HT_ * 2 + 2;   // accumulator doubles per lane (GW1, GW2, gb2)
  static constexpr int PARK = (OPT_ & MMA2_PARK_GW) ? ((NGW * 32 > NT_ * SCR1) ? NGW * 32 - NT_ * SCR1 : 0) : 0;   // beyond the scratch tiles it aliases
  static constexpr int PFL = 1 + KSU + 3 * Base::NX;       // prefetch slots (doubles) per lane and tile: time, data, 3 per covariate
  static constexpr int PFW = (OPT_ & 32) ? NT_ * 32 * PFL : 0;   // per warp
  static constexpr int WARP_FWD = PFW;
  static constexpr int RB = (OPT_ & MMA2_RING3) ? 3 : 2;   // stage-record slots per warp
  static constexpr int WARP_GRAD = RB * RECW + NT_ * SCR1 + PARK + PFW;
  static constexpr int NA = (NT_ == 1) ? HT_ : (HT_ >= 2 ? 2 : 1);   // accumulators per tile for K-long products
  static constexpr int TABN = (OPT_ & MMA2_TAB16) ? 64 * 16 : 64;
  static constexpr bool AFF = (OPT_ & 64) != 0 && D_ == 4 * KSU && KSU < Base::KS;
  static constexpr int NAFF = AFF ? (Base::NX + 1) * 8 * HT_ : 2;        // [bias | covariate columns][hidden]
};

// loads whose value is first needed an observation interval later: `volatile` pins the load where it is written (ptxas
// otherwise sinks it to the use and the prefetch becomes a stall: profiles/ncu_c3_mma2nt1_r02a.txt, long_sb on DSETP)
__device__ __forceinline__ double ld_pf(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
template <int OPT>
__device__ __forceinline__ double ld_pf_opt(const double* p) {
  if constexpr (OPT & 8) return *p; else return ld_pf(p);
}

// ---- staged prefetch: an 8-byte global -> shared copy now, an LDS when the value is needed -------------------------------
__device__ __forceinline__ void pf_issue(uint32_t saddr, const double* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void pf_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ double pf_read(uint32_t saddr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr) : "memory");
  return v;
}

// piecewise-linear covariate, referenced at the UPPER knot of the current interval:  x(t) = xhi + slope (t - vhi).
// Flat extrapolation = slope 0.  The description of the NEXT interval is requested when this one is entered: into the
// registers (p_t, p_x, p_inv) by default, into three shared-memory slots with MMA2_SMEM_PF (pf != 0).
struct CovLin {
  double vhi, xhi, slope;
  double p_t, p_x, p_inv;
  int j, kend;               // absolute index of the interval's lower knot (kbeg - 1 before the first knot), end of the knots
  uint32_t pf;               // shared-memory address of this lane's three prefetch slots (0: register prefetch)
};
constexpr double kCovBig = 1.0e300;

// prefetch the description of interval jn = [knot jn, knot jn + 1)
__device__ __forceinline__ void cov_prefetch(const DevProblem& P, CovLin& c, int jn) {
  const bool inner = jn + 1 < c.kend;
  const int ix = inner ? jn + 1 : c.kend - 1;
  if (c.pf) {
    pf_issue(c.pf, P.knot_t + ix); pf_issue(c.pf + 8, P.knot_x + ix); pf_issue(c.pf + 16, P.knot_inv + (inner ? jn : ix));
    return;
  }
  const double pt = ld_pf(P.knot_t + ix), px = ld_pf(P.knot_x + ix), pi = ld_pf(P.knot_inv + (inner ? jn : ix));
  c.p_t = inner ? pt : kCovBig;
  c.p_x = px;
  c.p_inv = inner ? pi : 0.0;
}

// one search per segment (task start): the interval that holds t
__device__ __forceinline__ void cov_init(const DevProblem& P, long long k0, int n, double t, CovLin& c, uint32_t pf) {
  c.j = 0; c.kend = 0; c.vhi = kCovBig; c.xhi = 0.0; c.slope = 0.0; c.p_t = kCovBig; c.p_x = 0.0; c.p_inv = 0.0; c.pf = pf;
  if (n <= 0) return;
  const int kbeg = (int)k0;
  c.kend = kbeg + n;
  const double* __restrict__ kt = P.knot_t;
  const double* __restrict__ kx = P.knot_x;
  if (t < kt[kbeg]) { c.j = kbeg - 1; c.vhi = kt[kbeg]; c.xhi = kx[kbeg]; c.slope = 0.0; }
  else if (t >= kt[c.kend - 1]) { c.j = c.kend - 1; c.vhi = kCovBig; c.xhi = kx[c.kend - 1]; c.slope = 0.0; return; }
  else {
    int lo = kbeg, hi = c.kend - 1;            // kt[lo] <= t < kt[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (kt[mid] <= t) lo = mid; else hi = mid; }
    c.j = lo; c.vhi = kt[lo + 1]; c.xhi = kx[lo + 1]; c.slope = (kx[lo + 1] - kx[lo]) * P.knot_inv[lo];
  }
  cov_prefetch(P, c, c.j + 1);
}

__device__ __forceinline__ double cov_value(const DevProblem& P, CovLin& c, double t) {
  while (t >= c.vhi) {                        // stage times never decrease inside a forward sweep
    if (c.pf) {                               // the request made when the current interval was entered has landed by now
      pf_wait();
      const bool inner = c.j + 2 < c.kend;    // the prefetched interval is j + 1
      const double pt = pf_read(c.pf), px = pf_read(c.pf + 8), pi = pf_read(c.pf + 16);
      c.p_t = inner ? pt : kCovBig; c.p_x = px; c.p_inv = inner ? pi : 0.0;
    }
    c.slope = (c.p_x - c.xhi) * c.p_inv; c.xhi = c.p_x; c.vhi = c.p_t;
    ++c.j;
    cov_prefetch(P, c, c.j + 1);
  }
  return fma(t - c.vhi, c.slope, c.xhi);
}

template <int NV, int TS, int V2>
__device__ __forceinline__ void tanh_vec_sel(const double (&zi)[NV], double (&ho)[NV], const double* s_tab, int toff) {
  if constexpr (V2 == 3) ude_tanh_fast3_vec<NV, TS>(zi, ho, s_tab, toff);
  else if constexpr (V2 == 2) ude_tanh_fast2_vec<NV, TS>(zi, ho, s_tab, toff);
  else ude_tanh_fast_vec<NV, TS>(zi, ho, s_tab, toff);
}

template <int N, int TB, int TS, int V2 = 1>
__device__ __forceinline__ void tanh_batches(double (&v)[N], const double* s_tab, int toff) {
  static_assert(N % 2 == 0 && (TB == 2 || TB == 4 || TB == 8), "batches of 2, 4 or 8 interleaved chains");
#pragma unroll
  for (int b = 0; b < N; b += TB) {
    if constexpr (TB == 8) {
      if (N - b >= 8) {
        double zi[8], ho[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) zi[q] = v[b + q < N ? b + q : 0];
        tanh_vec_sel<8, TS, V2>(zi, ho, s_tab, toff);
#pragma unroll
        for (int q = 0; q < 8; ++q) if (b + q < N) v[b + q] = ho[q];
        continue;
      }
    }
    if (TB >= 4 && N - b >= 4) {
#pragma unroll
      for (int b2 = b; b2 < (b + TB < N ? b + TB : N); b2 += 4) {
        double zi[4], ho[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) zi[q] = v[b2 + q < N ? b2 + q : 0];
        tanh_vec_sel<4, TS, V2>(zi, ho, s_tab, toff);
#pragma unroll
        for (int q = 0; q < 4; ++q) if (b2 + q < N) v[b2 + q] = ho[q];
      }
    } else {
#pragma unroll
      for (int b2 = b; b2 < (b + TB < N ? b + TB : N); b2 += 2) {
        double zi[2] = {v[b2], v[b2 + 1 < N ? b2 + 1 : 0]}, ho[2];
        tanh_vec_sel<2, TS, V2>(zi, ho, s_tab, toff);
        v[b2] = ho[0];
        if (b2 + 1 < N) v[b2 + 1] = ho[1];
      }
    }
  }
}

// ---- one RHS evaluation for the warp's NT x 8 segments; xs[tile] = this lane's input slots -------------------------
template <class MG>
__device__ __forceinline__ void mma2_stage_forward(const double* __restrict__ s_w, const double* s_tab, int lane,
                                                   const double (&xs)[MG::NT][MG::KS], const double (&b2own)[2],
                                                   double (&h)[MG::NT][MG::HT][2], double (&y)[MG::NT][2],
                                                   const double* s_aff, const double (&X)[MG::NT][MG::NXA]) {
  constexpr int NT = MG::NT, HT = MG::HT, KS = MG::AFF ? MG::KSU : MG::KS, NA = MG::NA;
  double z[NT][HT][2];
  if constexpr (MG::AFF) {
    // lane (r, c) owns hidden units 8t + 2c, 8t + 2c + 1 of segment r
#pragma unroll
    for (int t = 0; t < HT; ++t) {
      const double2 b = *reinterpret_cast<const double2*>(s_aff + 8 * t + 2 * (lane & 3));
#pragma unroll
      for (int tl = 0; tl < NT; ++tl) { z[tl][t][0] = b.x; z[tl][t][1] = b.y; }
#pragma unroll
      for (int q = 0; q < MG::NX; ++q) {
        const double2 w = *reinterpret_cast<const double2*>(s_aff + (q + 1) * 8 * HT + 8 * t + 2 * (lane & 3));
#pragma unroll
        for (int tl = 0; tl < NT; ++tl) { z[tl][t][0] = fma(X[tl][q], w.x, z[tl][t][0]); z[tl][t][1] = fma(X[tl][q], w.y, z[tl][t][1]); }
      }
    }
  } else {
#pragma unroll
  for (int tl = 0; tl < NT; ++tl)
#pragma unroll
    for (int t = 0; t < HT; ++t) { z[tl][t][0] = 0.0; z[tl][t][1] = 0.0; }
  }
#pragma unroll
  for (int e = 0; e < KS; ++e) {
#pragma unroll
    for (int t = 0; t < HT; ++t) {
      const double bf = s_w[(t * MG::KS + e) * 32 + lane];
#pragma unroll
      for (int tl = 0; tl < NT; ++tl) dmma(z[tl][t], xs[tl][e], bf);
    }
  }
  {
    double v[NT * HT * 2];
#pragma unroll
    for (int tl = 0; tl < NT; ++tl)
#pragma unroll
      for (int t = 0; t < HT; ++t) { v[(tl * HT + t) * 2] = z[tl][t][0]; v[(tl * HT + t) * 2 + 1] = z[tl][t][1]; }
    if constexpr (MG::OPT & MMA2_ABL_NO_TANH) {
#pragma unroll
      for (int i = 0; i < NT * HT * 2; ++i) v[i] = fma(v[i], 0.125, 0.25);
    } else
    tanh_batches<NT * HT * 2, MG::TB, (MG::OPT & MMA2_TAB16) ? 16 : 1, ((MG::OPT & MMA2_TANH3) ? 3 : (MG::OPT & MMA2_TANH2) ? 2 : 1)>(v, s_tab, (MG::OPT & MMA2_TAB16) ? (lane & 15) : 0);
#pragma unroll
    for (int tl = 0; tl < NT; ++tl)
#pragma unroll
      for (int t = 0; t < HT; ++t) { h[tl][t][0] = v[(tl * HT + t) * 2]; h[tl][t][1] = v[(tl * HT + t) * 2 + 1]; }
  }
  // layer 2: NA accumulators per tile (independent DMMA chains); b2 rides in the first one
  double yt[NT][NA][2];
#pragma unroll
  for (int tl = 0; tl < NT; ++tl)
#pragma unroll
    for (int a = 0; a < NA; ++a) { yt[tl][a][0] = a == 0 ? b2own[0] : 0.0; yt[tl][a][1] = a == 0 ? b2own[1] : 0.0; }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int t = 0; t < HT; ++t) {
      const double bf = s_w[(MG::NF1 + 2 * t + j) * 32 + lane];
#pragma unroll
      for (int tl = 0; tl < NT; ++tl) dmma(yt[tl][t % NA], h[tl][t][j], bf);
    }
  }
#pragma unroll
  for (int tl = 0; tl < NT; ++tl) {
    y[tl][0] = yt[tl][0][0]; y[tl][1] = yt[tl][0][1];
#pragma unroll
    for (int a = 1; a < NA; ++a) { y[tl][0] += yt[tl][a][0]; y[tl][1] += yt[tl][a][1]; }
  }
}

template <int D_, int NX_, int HT_, int NT_, int SOLVER, int LK, bool GRAD, int MINB, int OPT = 0, int TB = 4>
__global__ void __launch_bounds__(kBlockThreads, MINB) ude_mma2_kernel(const DevProblem P) {
  using MG = Mma2Geo<D_, NX_, HT_, NT_, OPT, TB>;
  using T = Tab<SOLVER>;
  static_assert(LK == LK_STATE || LK == LK_TRAJ, "derivative matching stays with ude_mma_kernel");
  constexpr int D = D_, NX = NX_, HT = HT_, NT = NT_, KS = MG::KS, KSU = MG::KSU, NXA = MG::NXA, NA = MG::NA;
  __shared__ double s_tab[MG::TABN];
  __shared__ __align__(8) uint64_t s_bar[kBlockThreads / 32][MG::RB];
  __shared__ __align__(16) double s_aff[MG::NAFF];
  extern __shared__ __align__(16) double s_dyn[];
  // [n_pm + 1 (block partial) padded even][weight fragments][per warp: 2 stage records (NT tiles each) | NT scratch tiles (+ park)]
  double* s_red = s_dyn;
  double* s_w = s_dyn + (P.n_pm + 1 + 1) / 2 * 2;
  constexpr int kPerWarp = GRAD ? MG::WARP_GRAD : MG::WARP_FWD;
  constexpr bool SPF = (OPT & MMA2_SMEM_PF) != 0;
  double* s_warp = s_w + MG::NFRAG * 32 + (size_t)(threadIdx.x >> 5) * kPerWarp;
  double* s_stage = s_warp;
  double* s_scr = s_warp + (GRAD ? MG::RB * MG::RECW : 0);
  // MMA2_SMEM_PF: this lane's prefetch slots, [tile][lane][time | data (KSU) | 3 per covariate]
  const uint32_t s_pf = SPF ? smem_u32(s_warp + (GRAD ? MG::RB * MG::RECW + NT * MG::SCR1 + MG::PARK : 0)) + (uint32_t)((threadIdx.x & 31) * MG::PFL * 8) : 0u;
  auto pf_slot = [&](int tl, int k) -> uint32_t { return s_pf + (uint32_t)((tl * 32 * MG::PFL + k) * 8); };
  uint64_t* bar = s_bar[threadIdx.x >> 5];

  const int lane = threadIdx.x & 31;
  const int r = lane >> 2, c = lane & 3;
  const long long warps_per_block = blockDim.x >> 5;
  const long long warp_global = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const long long total_warps = (long long)gridDim.x * warps_per_block;
  const int S = P.substeps;
  const double invS = 1.0 / (double)S;
  constexpr uint32_t kRecBytes = MG::RECW * 8;
  double* __restrict__ stash = GRAD ? reinterpret_cast<double*>(P.stash) + warp_global * P.stash_rows_per_warp * 32 : nullptr;

  if constexpr ((OPT & MMA2_TANH3) != 0) { if constexpr (OPT & MMA2_TAB16) ude_load_exp2_tabb_n<16>(s_tab); else ude_load_exp2_tabb_n<1>(s_tab); }
  else if constexpr (OPT & MMA2_TAB16) ude_load_exp2_tab_n<16>(s_tab); else ude_load_exp2_tab(s_tab);
  for (long long k = threadIdx.x; k <= P.n_pm; k += blockDim.x) s_red[k] = 0.0;
  const double* __restrict__ pm = P.theta + (long long)D * P.model_col0[1];        // single model: [uhat | pm]
  build_frags<typename MG::Base>(P, pm, s_w);
  if constexpr (MG::AFF)
    for (int k = threadIdx.x; k < MG::NAFF; k += blockDim.x) {
      const int q = k / (8 * HT), j = k % (8 * HT);             // q = 0: b1 (slot D + NX), q >= 1: covariate q - 1 (slot D + q - 1)
      s_aff[k] = w1_slot<typename MG::Base>(P, pm, j, q == 0 ? D + NX : D + q - 1);
    }
  if (GRAD && lane == 0) { for (int b = 0; b < MG::RB; ++b) mbar_init(&bar[b], 1); }
  uint32_t bar_phase = 0;
  __syncthreads();

  double b2own[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) b2own[j] = (c + 4 * j < D) ? pm[P.off_b2 + c + 4 * j] : 0.0;

  // gradient accumulators (C fragments) -- summed over the warp's segments by the MMA itself
  double GW1[MG::MT][HT][2], GW2[HT][2], gb2[2] = {0.0, 0.0};
#pragma unroll
  for (int t = 0; t < HT; ++t) {
    GW2[t][0] = GW2[t][1] = 0.0;
#pragma unroll
    for (int m = 0; m < MG::MT; ++m) GW1[m][t][0] = GW1[m][t][1] = 0.0;
  }
  double loss = 0.0;
  // MMA2_PARK_GW: the accumulators live in the (forward-idle) scratch tiles between reverse sweeps, lane-major
  double* s_park = s_scr;
  auto park_store = [&]() {
    if constexpr (GRAD && (OPT & MMA2_PARK_GW)) {
      int q = 0;
#pragma unroll
      for (int t = 0; t < HT; ++t) {
#pragma unroll
        for (int m = 0; m < MG::MT; ++m) { s_park[(q++) * 32 + lane] = GW1[m][t][0]; s_park[(q++) * 32 + lane] = GW1[m][t][1]; }
        s_park[(q++) * 32 + lane] = GW2[t][0]; s_park[(q++) * 32 + lane] = GW2[t][1];
      }
      s_park[(q++) * 32 + lane] = gb2[0]; s_park[(q++) * 32 + lane] = gb2[1];
      __syncwarp();
    }
  };
  auto park_load = [&]() {
    if constexpr (GRAD && (OPT & MMA2_PARK_GW)) {
      int q = 0;
#pragma unroll
      for (int t = 0; t < HT; ++t) {
#pragma unroll
        for (int m = 0; m < MG::MT; ++m) { GW1[m][t][0] = s_park[(q++) * 32 + lane]; GW1[m][t][1] = s_park[(q++) * 32 + lane]; }
        GW2[t][0] = s_park[(q++) * 32 + lane]; GW2[t][1] = s_park[(q++) * 32 + lane];
      }
      gb2[0] = s_park[(q++) * 32 + lane]; gb2[1] = s_park[(q++) * 32 + lane];
      __syncwarp();
    }
  };
  park_store();

  const double* __restrict__ uh = P.uhat_x ? P.uhat_x : P.theta;        // single model: uhat is indexed by global column
  double* __restrict__ guh = P.guhat_x ? P.guhat_x : P.grad;
  const bool multi = P.loss_kind == UDE_LOSS_MULTI_SHOOT;

  // this lane's input slots from its state slots, the covariates and the constant 1 that carries b1
  auto make_xs = [&](const double (&U)[KSU], const double (&X)[NXA], double (&xs)[KS]) {
#pragma unroll
    for (int e = 0; e < KS; ++e) {
      const int s = c + 4 * e;
      double v = 0.0;
      if (e < KSU) { if (s < D) v = U[e < KSU ? e : 0]; }
      if (s == D + NX) v = 1.0;
#pragma unroll
      for (int q = 0; q < NX; ++q) if (s == D + q) v = X[q];
      xs[e] = v;
    }
  };

  Seg nxt[NT];
#pragma unroll
  for (int tl = 0; tl < NT; ++tl) {
    const long long t0 = P.task_begin + warp_global;
    nxt[tl] = P.segs[(t0 < P.n_tasks ? t0 : P.n_tasks - 1) * (8 * NT) + 8 * tl + r];
  }

  for (long long task = P.task_begin + warp_global; task < P.n_tasks; task += total_warps) {
    Seg sg[NT];
#pragma unroll
    for (int tl = 0; tl < NT; ++tl) sg[tl] = nxt[tl];
    {                                             // segment table entries of my next task: loaded a whole task ahead
      const long long tn = task + total_warps;
#pragma unroll
      for (int tl = 0; tl < NT; ++tl) nxt[tl] = P.segs[(tn < P.n_tasks ? tn : task) * (8 * NT) + 8 * tl + r];
    }
    bool is_null[NT], has_prev[NT];
    int c0[NT], L[NT];
    double wpt[NT];
    int maxL = 0, pre = 0, first_scored = 0;
    bool from_theta = true;
    double eps = 0.0;
#pragma unroll
    for (int tl = 0; tl < NT; ++tl) {
      is_null[tl] = (sg[tl].flags & SEG_NULL) != 0;
      has_prev[tl] = true; wpt[tl] = 0.0;
      if (LK == LK_STATE) {
        const bool skip = is_null[tl] || (sg[tl].flags & SEG_SKIP_PROC);
        has_prev[tl] = !is_null[tl] && !(sg[tl].flags & SEG_NO_PREV);
        c0[tl] = has_prev[tl] ? sg[tl].col - 1 : sg[tl].col;
        L[tl] = (skip || !has_prev[tl]) ? 0 : 1;
      } else {
        c0[tl] = sg[tl].col;
        L[tl] = is_null[tl] ? 0 : sg[tl].L;
        wpt[tl] = is_null[tl] ? 0.0 : (multi ? 1.0 / ((double)D * (double)P.lengths[sg[tl].series]) : P.shoot_weight[0]);
        maxL = max(maxL, L[tl]);
      }
    }
    if (LK == LK_STATE) maxL = 1;
    else {
      maxL = (int)__reduce_max_sync(0xffffffffu, (unsigned)maxL);
      from_theta = multi || P.shoot_from_theta;
      first_scored = multi ? 0 : P.shoot_first_scored;
      eps = multi ? P.preroll_eps : 0.0;
      pre = eps > 0.0 ? 1 : 0;
    }
    const int nsteps = pre + maxL * S;

    // ---- everything the first interval needs, in one batch of independent loads ------------------------------------
    double u[NT][KSU], tq0[NT], tq1[NT], dq[NT][KSU];
#pragma unroll
    for (int tl = 0; tl < NT; ++tl) {
      tq0[tl] = P.times[c0[tl]];
      if constexpr (SPF) { pf_issue(pf_slot(tl, 0), P.times + c0[tl] + min(1, L[tl])); tq1[tl] = 0.0; }
      else tq1[tl] = P.times[c0[tl] + min(1, L[tl])];
#pragma unroll
      for (int e = 0; e < KSU; ++e) {
        const int s = c + 4 * e;
        u[tl][e] = (s < D) ? (from_theta ? uh[(long long)D * c0[tl] + s] : P.data[(long long)D * c0[tl] + s]) : 0.0;
        dq[tl][e] = 0.0;
        if (s < D && LK == LK_TRAJ) {
          if constexpr (SPF) pf_issue(pf_slot(tl, 1 + e), P.data + (long long)D * c0[tl] + s);
          else dq[tl][e] = P.data[(long long)D * c0[tl] + s];
        }
      }
    }
    CovLin cov[NT][NXA];
    if (NX > 0) {
#pragma unroll
      for (int tl = 0; tl < NT; ++tl)
#pragma unroll
        for (int v = 0; v < NX; ++v) {
          const long long k0 = P.knot_off[(long long)sg[tl].series * NX + v];
          const int n = (int)(P.knot_off[(long long)sg[tl].series * NX + v + 1] - k0);
          cov_init(P, k0, n, tq0[tl] - eps, cov[tl][v], SPF ? pf_slot(tl, 1 + KSU + 3 * v) : 0u);
        }
    }

    // ------------------------------------------ forward sweep ------------------------------------------------------
    {
      int i = pre ? -1 : 0, n = pre ? S - 1 : 0;
      double t0[NT], hs[NT], nd = 0.0;
#pragma unroll
      for (int tl = 0; tl < NT; ++tl) { t0[tl] = tq0[tl] - eps; hs[tl] = is_null[tl] ? 0.0 : eps; }
      for (int k = 0; k < nsteps; ++k) {
        if (k >= pre && n == 0) {
          if constexpr (SPF) {                 // what the previous event (or the task start) requested has landed
            pf_wait();
#pragma unroll
            for (int tl = 0; tl < NT; ++tl) {
              tq1[tl] = pf_read(pf_slot(tl, 0));
              if (LK == LK_TRAJ) {
#pragma unroll
                for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) dq[tl][e] = pf_read(pf_slot(tl, 1 + e));
              }
            }
          }
#pragma unroll
          for (int tl = 0; tl < NT; ++tl) {
            if (LK == LK_TRAJ) {
              const double w = (i <= L[tl] && i >= first_scored) ? wpt[tl] : 0.0;
#pragma unroll
              for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) {
                const double res = dq[tl][e] - u[tl][e];
                loss = fma(w * res, res, loss);
                if (!GRAD && P.traj_out != nullptr && i <= L[tl] && !is_null[tl]) P.traj_out[(long long)D * (c0[tl] + i) + c + 4 * e] = u[tl][e];
              }
            }
            t0[tl] = tq0[tl];
            hs[tl] = (i < L[tl]) ? (tq1[tl] - tq0[tl]) * invS : 0.0;
            // loads for the NEXT interval: consumed a whole interval (S steps) from now
            tq0[tl] = tq1[tl];
            if constexpr (SPF) pf_issue(pf_slot(tl, 0), P.times + c0[tl] + min(i + 2, L[tl]));
            else tq1[tl] = ld_pf_opt<OPT>(P.times + c0[tl] + min(i + 2, L[tl]));
            if (LK == LK_TRAJ) {
#pragma unroll
              for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) {
                const double* src = P.data + (long long)D * (c0[tl] + min(i + 1, L[tl])) + c + 4 * e;
                if constexpr (SPF) pf_issue(pf_slot(tl, 1 + e), src); else dq[tl][e] = ld_pf_opt<OPT>(src);
              }
            }
          }
          nd = 0.0;
        }
        // one explicit RK step on this lane's state slots
        double kst[NT][T::S][KSU], acc[NT][KSU], X[NT][NXA];
#pragma unroll
        for (int tl = 0; tl < NT; ++tl) {
#pragma unroll
          for (int e = 0; e < KSU; ++e) acc[tl][e] = 0.0;
#pragma unroll
          for (int v = 0; v < NXA; ++v) X[tl][v] = 0.0;
        }
#pragma unroll
        for (int st = 0; st < T::S; ++st) {
          double xs[NT][KS], h[NT][HT][2], y[NT][2];
#pragma unroll
          for (int tl = 0; tl < NT; ++tl) {
            double U[KSU];
#pragma unroll
            for (int e = 0; e < KSU; ++e) {
              double a = 0.0;
#pragma unroll
              for (int l = 0; l < st; ++l) if (T::a(st, l) != 0.0) a = fma(T::a(st, l), kst[tl][l][e], a);
              U[e] = (st == 0) ? u[tl][e] : fma(hs[tl], a, u[tl][e]);
            }
            if (NX > 0 && !(OPT & MMA2_ABL_NO_COV) && (st == 0 || T::c(st) != T::c(st > 0 ? st - 1 : 0))) {
              const double ts = fma(nd, hs[tl], t0[tl]) + T::c(st) * hs[tl];
#pragma unroll
              for (int v = 0; v < NX; ++v) X[tl][v] = cov_value(P, cov[tl][v], ts);
            }
            make_xs(U, X[tl], xs[tl]);
          }
          mma2_stage_forward<MG>(s_w, s_tab, lane, xs, b2own, h, y, s_aff, X);
          if (GRAD && (OPT & MMA2_DIRECT_ST)) {
            // stage record straight from the registers to the stash, in the layout the reverse sweep bulk-loads
            double* __restrict__ grec = stash + (long long)(k * T::S + st) * MG::RECW;
#pragma unroll
            for (int tl = 0; tl < NT; ++tl) {
              double* tb = grec + tl * MG::REC1;
#pragma unroll
              for (int t2 = 0; t2 < HT; ++t2) { tb[MG::hpos(r, 8 * t2 + c)] = h[tl][t2][0]; tb[MG::hpos(r, 8 * t2 + c + 4)] = h[tl][t2][1]; }
#pragma unroll
              for (int e = 0; e < KS; ++e) tb[8 * MG::HS + MG::xpos(r, c + 4 * e)] = xs[tl][e];
            }
          } else
          if (GRAD) {
            // stage record -> shared memory -> TMA bulk store to the stash (two-buffer ring, one record per stage)
            const int rec = k * T::S + st;
            double* sbuf = s_stage + (rec & 1) * MG::RECW;
            if (lane == 0 && !(OPT & MMA2_ABL_NO_TMA)) bulk_wait_read<1>();
            __syncwarp();
            if constexpr (!(OPT & MMA2_ABL_NO_REC))
#pragma unroll
            for (int tl = 0; tl < NT; ++tl) {
              double* tb = sbuf + tl * MG::REC1;
#pragma unroll
              for (int t2 = 0; t2 < HT; ++t2) { tb[MG::hpos(r, 8 * t2 + c)] = h[tl][t2][0]; tb[MG::hpos(r, 8 * t2 + c + 4)] = h[tl][t2][1]; }
#pragma unroll
              for (int e = 0; e < KS; ++e) tb[8 * MG::HS + MG::xpos(r, c + 4 * e)] = xs[tl][e];
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0 && !(OPT & MMA2_ABL_NO_TMA)) { bulk_store(stash + (long long)rec * MG::RECW, sbuf, kRecBytes); bulk_commit(); }
          }
#pragma unroll
          for (int tl = 0; tl < NT; ++tl)
#pragma unroll
            for (int e = 0; e < KSU; ++e) {
              const double du = (c + 4 * e < D && e < 2) ? y[tl][e < 2 ? e : 0] : 0.0;
              kst[tl][st][e] = du; acc[tl][e] = fma(T::b(st), du, acc[tl][e]);
            }
        }
#pragma unroll
        for (int tl = 0; tl < NT; ++tl)
#pragma unroll
          for (int e = 0; e < KSU; ++e) u[tl][e] = fma(hs[tl], acc[tl][e], u[tl][e]);
        nd += 1.0;
        if (++n == S) { n = 0; ++i; }
      }
      if (GRAD && !(OPT & MMA2_ABL_NO_TMA)) {
        if constexpr (OPT & MMA2_DIRECT_ST) {       // generic-proxy stores of every lane -> async-proxy (TMA) loads by lane 0
          asm volatile("fence.proxy.async;" ::: "memory");
          __syncwarp();
        }
        if (lane == 0) {
          if constexpr (!(OPT & MMA2_DIRECT_ST)) bulk_wait_all<0>();
          if (nsteps > 0) {
            const int rec = nsteps * T::S - 1, b = (MG::RB == 3) ? 0 : (rec & 1);      // three slots: the reverse sweep starts in slot 0
            mbar_expect_tx(&bar[b], kRecBytes);
            bulk_load(s_stage + b * MG::RECW, stash + (long long)rec * MG::RECW, kRecBytes, &bar[b]);
            if (MG::RB == 3 && rec > 0) {
              mbar_expect_tx(&bar[1], kRecBytes);
              bulk_load(s_stage + MG::RECW, stash + (long long)(rec - 1) * MG::RECW, kRecBytes, &bar[1]);
            }
          }
        }
        __syncwarp();
      }
    }
    // ------------------------------------------ loss at the last point ------------------------------------------------
    if constexpr (SPF && LK == LK_TRAJ) {      // data at point min(maxL, L): requested by the last event (or at task start)
      pf_wait();
#pragma unroll
      for (int tl = 0; tl < NT; ++tl)
#pragma unroll
        for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) dq[tl][e] = pf_read(pf_slot(tl, 1 + e));
    } else if constexpr (SPF) pf_wait();
    double lam[NT][KSU], gc[NT][KSU], gp[NT][KSU];
    // reverse-sweep prefetch registers: times of the interval being reversed, data column at its lower end
    double rt_lo[NT], rt_hi[NT], dr[NT][KSU];
#pragma unroll
    for (int tl = 0; tl < NT; ++tl) {
#pragma unroll
      for (int e = 0; e < KSU; ++e) { lam[tl][e] = 0.0; gc[tl][e] = 0.0; gp[tl][e] = 0.0; dr[tl][e] = 0.0; }
      if (GRAD) {
        const int ir = max(maxL - 1, 0);
        rt_lo[tl] = P.times[c0[tl] + min(ir, L[tl])];
        rt_hi[tl] = P.times[c0[tl] + min(ir + 1, L[tl])];
      } else { rt_lo[tl] = 0.0; rt_hi[tl] = 0.0; }
      if (LK == LK_STATE) {
        const int col = sg[tl].col;
        const double dt = P.times[col] - P.times[c0[tl]];
        const double wsc = (L[tl] == 1) ? P.proc_scale[0] * (P.proc_inv_dt ? 1.0 / dt : 1.0) : 0.0;
        const double os = is_null[tl] ? 0.0 : P.obs_scale[0];
        const bool obs_first = (sg[tl].flags & SEG_OBS_FIRST) && has_prev[tl];
        // residual vectors are spread over the quad: gather them (once per segment) for the d x d quadratic forms
        double rown[2], eown[2], fown[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const bool st = (c + 4 * j) < D && j < KSU;
          rown[j] = st ? uh[(long long)D * col + c + 4 * j] - u[tl][j < KSU ? j : 0] : 0.0;
          eown[j] = st ? P.data[(long long)D * col + c + 4 * j] - uh[(long long)D * col + c + 4 * j] : 0.0;
          fown[j] = (st && obs_first) ? P.data[(long long)D * c0[tl] + c + 4 * j] - uh[(long long)D * c0[tl] + c + 4 * j] : 0.0;
        }
        double rf[D], ef[D], ff[D], sym[D], es[D], fs[D];
        const int qb = lane & ~3;
#pragma unroll
        for (int b = 0; b < D; ++b) {
          rf[b] = __shfl_sync(0xffffffffu, rown[b >> 2], qb | (b & 3));
          ef[b] = __shfl_sync(0xffffffffu, eown[b >> 2], qb | (b & 3));
          ff[b] = __shfl_sync(0xffffffffu, fown[b >> 2], qb | (b & 3));
        }
        const double q1 = quad_form<D>(P.B, P.Bsym, rf, sym);
        const double q2 = quad_form<D>(P.A, P.Asym, ef, es);
        const double q3 = quad_form<D>(P.A, P.Asym, ff, fs);
        if (c == 0) loss += wsc * q1 + os * (q2 + q3);
#pragma unroll
        for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) {
#pragma unroll
          for (int b = 0; b < D; ++b) if (b == c + 4 * e) {
            lam[tl][e] = -wsc * sym[b]; gc[tl][e] = wsc * sym[b] - os * es[b]; gp[tl][e] = -os * fs[b];
          }
        }
        if (!GRAD && P.traj_out != nullptr && !is_null[tl]) {
#pragma unroll
          for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) {
            if (L[tl] == 1) P.traj_out[(long long)D * col + c + 4 * e] = u[tl][e];
            if (!has_prev[tl] || (sg[tl].flags & SEG_OBS_FIRST)) P.traj_out[(long long)D * c0[tl] + c + 4 * e] = uh[(long long)D * c0[tl] + c + 4 * e];
          }
        }
      } else {
        const double w = (maxL <= L[tl] && maxL >= first_scored) ? wpt[tl] : 0.0;
#pragma unroll
        for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) {
          const double res = dq[tl][e] - u[tl][e];            // dq = data at point min(maxL, L)
          loss = fma(w * res, res, loss);
          lam[tl][e] = -2.0 * w * res;
          if (!GRAD && P.traj_out != nullptr && maxL <= L[tl] && !is_null[tl] && !(multi && L[tl] == P.pred_length))
            P.traj_out[(long long)D * (c0[tl] + L[tl]) + c + 4 * e] = u[tl][e];
        }
      }
    }
    // ------------------------------------------ reverse sweep ----------------------------------------------------------
    if (GRAD) {
      park_load();
      int i = maxL - 1, n = S - 1;
      double hs[NT];
#pragma unroll
      for (int tl = 0; tl < NT; ++tl) hs[tl] = 0.0;
      [[maybe_unused]] int rb3 = 0;            // MMA2_RING3: slot of the record being reversed (0, 1, 2, 0, ...)
      for (int k = nsteps - 1; k >= 0; --k) {
        if (k >= pre) {
          if (n == S - 1) {
            if constexpr (SPF) {
              if (i != maxL - 1) {             // the lower time of this interval was requested when the one above started
                pf_wait();
#pragma unroll
                for (int tl = 0; tl < NT; ++tl) { rt_hi[tl] = rt_lo[tl]; rt_lo[tl] = pf_read(pf_slot(tl, 0)); }
              }
            }
#pragma unroll
            for (int tl = 0; tl < NT; ++tl) {
              hs[tl] = (i < L[tl]) ? (rt_hi[tl] - rt_lo[tl]) * invS : 0.0;
              // loads for the interval below and for this interval's lower-end data column (both used S steps from now)
              if (LK == LK_TRAJ) {
#pragma unroll
                for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) {
                  const double* src = P.data + (long long)D * (c0[tl] + min(i, L[tl])) + c + 4 * e;
                  if constexpr (SPF) pf_issue(pf_slot(tl, 1 + e), src); else dr[tl][e] = ld_pf_opt<OPT>(src);
                }
              }
              if constexpr (SPF) pf_issue(pf_slot(tl, 0), P.times + c0[tl] + min(max(i - 1, 0), L[tl]));
              else { rt_hi[tl] = rt_lo[tl]; rt_lo[tl] = ld_pf_opt<OPT>(P.times + c0[tl] + min(max(i - 1, 0), L[tl])); }
            }
          }
        } else {
#pragma unroll
          for (int tl = 0; tl < NT; ++tl) hs[tl] = is_null[tl] ? 0.0 : eps;
        }
        double Ub[NT][T::S][KSU], lamacc[NT][KSU], U0[NT][KSU];
#pragma unroll
        for (int tl = 0; tl < NT; ++tl)
#pragma unroll
          for (int e = 0; e < KSU; ++e) { lamacc[tl][e] = 0.0; U0[tl][e] = 0.0; }
#pragma unroll
        for (int ii = 0; ii < T::S; ++ii) {
          const int st = T::S - 1 - ii;
          const int rec = k * T::S + st;
          int b;
          if constexpr (MG::RB == 3) { b = rb3; rb3 = (rb3 == 2) ? 0 : rb3 + 1; } else b = rec & 1;
          if constexpr (!(OPT & MMA2_ABL_NO_TMA)) {
          if constexpr (MG::RB == 3) {
            if (rec > 1 && lane == 0) {        // the record two stages ahead goes into the slot the previous stage has just released
              const int b2 = (b == 0) ? 2 : b - 1;
              mbar_expect_tx(&bar[b2], kRecBytes);
              bulk_load(s_stage + b2 * MG::RECW, stash + (long long)(rec - 2) * MG::RECW, kRecBytes, &bar[b2]);
            }
          } else
          if (rec > 0 && lane == 0) {          // prefetch the next record to be reversed into the other buffer
            mbar_expect_tx(&bar[b ^ 1], kRecBytes);
            bulk_load(s_stage + (b ^ 1) * MG::RECW, stash + (long long)(rec - 1) * MG::RECW, kRecBytes, &bar[b ^ 1]);
          }
          mbar_wait(&bar[b], (bar_phase >> b) & 1u);
          bar_phase ^= (1u << b);
          }
          const double* __restrict__ sbuf = s_stage + b * MG::RECW;
          double h[NT][HT][2], kb[NT][2];
#pragma unroll
          for (int tl = 0; tl < NT; ++tl) {
            const double* tb = sbuf + tl * MG::REC1;
#pragma unroll
            for (int t = 0; t < HT; ++t) { h[tl][t][0] = tb[MG::hpos(r, 8 * t + c)]; h[tl][t][1] = tb[MG::hpos(r, 8 * t + c + 4)]; }
            if (st == 0) {
#pragma unroll
              for (int e = 0; e < KSU; ++e) U0[tl][e] = tb[8 * MG::HS + MG::xpos(r, c + 4 * e)];
            }
            // k-bar on my output slots (outputs o = c + 4j coincide with state slots e = j)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              double a = 0.0;
              if (j < KSU) {
                a = T::b(st) * lam[tl][j < KSU ? j : 0];
#pragma unroll
                for (int l = st + 1; l < T::S; ++l) if (T::a(l, st) != 0.0) a = fma(T::a(l, st), Ub[tl][l][j < KSU ? j : 0], a);
              }
              kb[tl][j] = ((c + 4 * j) < D) ? hs[tl] * a : 0.0;
            }
          }
          double zb[NT][HT][2];
#pragma unroll
          for (int tl = 0; tl < NT; ++tl)
#pragma unroll
            for (int t = 0; t < HT; ++t) { zb[tl][t][0] = zb[tl][t][1] = 0.0; }
#pragma unroll
          for (int e = 0; e < MG::KO; ++e) {
#pragma unroll
            for (int t = 0; t < HT; ++t) {
              const double bf = s_w[(MG::NF1 + MG::NF2 + t * MG::KO + e) * 32 + lane];
#pragma unroll
              for (int tl = 0; tl < NT; ++tl) dmma(zb[tl][t], kb[tl][e], bf);
            }
          }
#pragma unroll
          for (int tl = 0; tl < NT; ++tl) {
            double* s_zs = s_scr + tl * MG::SCR1;
            double* s_ks = s_zs + 8 * MG::HS;
#pragma unroll
            for (int t = 0; t < HT; ++t) {
              zb[tl][t][0] *= fma(-h[tl][t][0], h[tl][t][0], 1.0);
              zb[tl][t][1] *= fma(-h[tl][t][1], h[tl][t][1], 1.0);
              s_zs[MG::hpos(r, 8 * t + c)] = zb[tl][t][0]; s_zs[MG::hpos(r, 8 * t + c + 4)] = zb[tl][t][1];
            }
            s_ks[MG::kpos(r, c)] = kb[tl][0]; s_ks[MG::kpos(r, c + 4)] = kb[tl][1];
            gb2[0] += kb[tl][0]; gb2[1] += kb[tl][1];
          }
          // x-bar = W1^T z-bar on the state slots
          double xt[NT][NA][2];
#pragma unroll
          for (int tl = 0; tl < NT; ++tl)
#pragma unroll
            for (int a = 0; a < NA; ++a) { xt[tl][a][0] = 0.0; xt[tl][a][1] = 0.0; }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int t = 0; t < HT; ++t) {
              const double bf = s_w[(MG::NF1 + MG::NF2 + MG::NF3 + 2 * t + j) * 32 + lane];
#pragma unroll
              for (int tl = 0; tl < NT; ++tl) dmma(xt[tl][t % NA], zb[tl][t][j], bf);
            }
          }
#pragma unroll
          for (int tl = 0; tl < NT; ++tl) {
            double xsum[2] = {xt[tl][0][0], xt[tl][0][1]};
#pragma unroll
            for (int a = 1; a < NA; ++a) { xsum[0] += xt[tl][a][0]; xsum[1] += xt[tl][a][1]; }
#pragma unroll
            for (int e = 0; e < KSU; ++e) {
              const double ub = (e < 2 && c + 4 * e < D) ? xsum[e < 2 ? e : 0] : 0.0;
              Ub[tl][st][e] = ub; lamacc[tl][e] += ub;
            }
          }
          __syncwarp();                       // z-bar / k-bar tiles complete
          // weight gradients: contraction over the warp's segments, four at a time
          if constexpr (!(OPT & MMA2_ABL_NO_GW))
#pragma unroll
          for (int tl = 0; tl < NT; ++tl) {
            const double* tb = sbuf + tl * MG::REC1;
            const double* s_zs = s_scr + tl * MG::SCR1;
            const double* s_ks = s_zs + 8 * MG::HS;
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
              const int seg = 4 * s2 + c;
              double aX[MG::MT];
#pragma unroll
              for (int m = 0; m < MG::MT; ++m) aX[m] = tb[8 * MG::HS + MG::xpos(seg, r + 8 * m)];
              const double aK = s_ks[MG::kpos(seg, r)];
#pragma unroll
              for (int t = 0; t < HT; ++t) {
                const double bZ = s_zs[MG::hpos(seg, 8 * t + r)];
                const double bH = tb[MG::hpos(seg, 8 * t + r)];
#pragma unroll
                for (int m = 0; m < MG::MT; ++m) dmma(GW1[m][t], aX[m], bZ);
                dmma(GW2[t], aK, bH);
              }
            }
          }
          __syncwarp();                       // everyone is done with buffer b and the scratch tiles
          // the record is dead: drop its (dirty) L2 lines instead of letting them be written back to HBM
          if (P.stash_discard && !(OPT & MMA2_ABL_NO_TMA)) {
            for (int ln = lane; ln < (int)(kRecBytes / 128); ln += 32)
              asm volatile("discard.global.L2 [%0], 128;" ::"l"(stash + (long long)rec * MG::RECW + ln * 16) : "memory");
          }
        }
#pragma unroll
        for (int tl = 0; tl < NT; ++tl)
#pragma unroll
          for (int e = 0; e < KSU; ++e) lam[tl][e] += lamacc[tl][e];
        if (k >= pre) {
          if (n == 0) {
            if (LK == LK_TRAJ) {
              if constexpr (SPF) {             // this interval's lower-end data column, requested S steps ago
                pf_wait();
#pragma unroll
                for (int tl = 0; tl < NT; ++tl)
#pragma unroll
                  for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) dr[tl][e] = pf_read(pf_slot(tl, 1 + e));
              }
#pragma unroll
              for (int tl = 0; tl < NT; ++tl) {
                const double w = (i <= L[tl] && i >= first_scored) ? wpt[tl] : 0.0;
#pragma unroll
                for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) lam[tl][e] = fma(-2.0 * w, dr[tl][e] - U0[tl][e], lam[tl][e]);
              }
            }
            --i; n = S - 1;
          } else --n;
        }
      }
#pragma unroll
      for (int tl = 0; tl < NT; ++tl) {
        if (!is_null[tl]) {
#pragma unroll
          for (int e = 0; e < KSU; ++e) if (c + 4 * e < D) {
            if (LK == LK_STATE) {
              atomicAdd(guh + (long long)D * sg[tl].col + c + 4 * e, gc[tl][e]);
              if (has_prev[tl]) atomicAdd(guh + (long long)D * c0[tl] + c + 4 * e, gp[tl][e] + lam[tl][e]);
            } else if (from_theta) {
              if (P.guhat_x) guh[(long long)D * c0[tl] + c + 4 * e] = lam[tl][e]; else atomicAdd(guh + (long long)D * c0[tl] + c + 4 * e, lam[tl][e]);
            }
          }
        }
      }
      park_store();
    }
  }
  park_load();

  // ---- flush: the C fragments already hold sums over the warp's segments; warps add into the block row in order ----
  {
    double l = loss;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) { gb2[0] += __shfl_xor_sync(0xffffffffu, gb2[0], off); gb2[1] += __shfl_xor_sync(0xffffffffu, gb2[1], off); }
    const int nw = blockDim.x >> 5, wib = threadIdx.x >> 5;
    for (int w = 0; w < nw; ++w) {
      if (wib == w) {
        if (GRAD) {
#pragma unroll
          for (int t = 0; t < HT; ++t) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int ncol = 2 * c + j;              // tile column n holds the hidden unit of lane (n & 3), value (n >> 2)
              const int hid = 8 * t + 2 * (ncol & 3) + (ncol >> 2);
              if (hid < P.H) {
#pragma unroll
                for (int m = 0; m < MG::MT; ++m) {
                  const int s = r + 8 * m;            // slot: input column, or the bias
                  if (s < D + NX) s_red[P.off_w1 + hid + (long long)P.H * s] += GW1[m][t][j];
                  else if (s == D + NX) s_red[P.off_b1 + hid] += GW1[m][t][j];
                }
                if (r < D) s_red[P.off_w2 + r + (long long)P.nn_out * hid] += GW2[t][j];
              }
            }
          }
          if (r == 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j) if (c + 4 * j < D) s_red[P.off_b2 + c + 4 * j] += gb2[j];
          }
        }
        if (lane == 0) s_red[P.n_pm] += l;
      }
      __syncthreads();
    }
    double* row = P.block_part + (long long)blockIdx.x * (P.n_pm + 1);
    for (long long k = threadIdx.x; k <= P.n_pm; k += blockDim.x) row[k] = s_red[k];
  }
}

}  // namespace ude
