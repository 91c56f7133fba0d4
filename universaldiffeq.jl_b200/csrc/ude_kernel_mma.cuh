// Tensor-core variant of the hot path for NODE-type right-hand sides (du = NN([u; X(t)])): the per-stage MLP of EIGHT
// segments per warp is evaluated with FP64 tensor-core MMAs (mma.sync.m8n8k4.f64, SASS DMMA).  tcgen05/UMMA has no
// FP64 kind, so DMMA is the only tensor-core route for the FP64 parity bar; on B200 it runs at the FP64 pipe's full
// rate (37.1 TFLOP/s measured, profiles/microbench_r01.json) while needing 1/8 of the issue slots of DFMA and -- the
// point of this kernel -- NO warp shuffles and only ~1/4 of the registers of the lane-per-hidden-unit kernel
// (ude_kernel.cuh), so three times as many segments are in flight per SM.
//
// Fragment algebra (m8n8k4: A 8x4 row-major, lane holds A[l>>2][l&3]; B 4x8, lane holds B[l&3][l>>2];
// C 8x8, lane holds C[l>>2][2(l&3)], C[l>>2][2(l&3)+1]).  With r = l>>2 (segment of the warp) and c = l&3:
//   * lane (r,c) OWNS the state / input slots {c, c+4, c+8..} of segment r and the hidden units {8t+2c, 8t+2c+1};
//   * every contraction index is permuted so that the values a lane produced (C layout) are exactly the values it must
//     supply as the next A operand: layer 1 takes k-step e = slots {4e..4e+3} -> lane supplies its own slot c+4e;
//     layer 2 / the x-bar product take k-step (t,j) = hidden {8t+2k+j} -> lane supplies its own h or z-bar; the
//     weights are pre-permuted once per block into B fragments in shared memory (one conflict-free LDS.64 each);
//   * the output column n = 2c+j of an 8-wide N tile is mapped to output / slot c+4j, so results land in the owner;
//   * the bias b1 rides along as an input slot that is constantly 1, b2 is added by the owner lane;
//   * only the weight-gradient contractions run over SEGMENTS (sum over the batch), which needs the operands
//     transposed (segment index on l&3): z-bar and k-bar go through a small per-warp shared-memory tile, h and x are
//     read straight from the stage record that the TMA brought back from the stash.
// Everything else (segment table, flat step loops, covariate cache, TMA-staged stash, fixed-order reduction) is shared
// with ude_kernel.cuh.
#pragma once
#include "ude_kernel.cuh"

namespace ude {

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int D_, int NX_, int HT_, int HW_ = 1, int TB_ = 4, int BX_ = 0>
struct MmaGeo {
  static constexpr int TB = TB_;                         // tanh evaluations interleaved per batch (4 or 8)
  // BX: the bias b1 stays OUT of the MMAs.  When D + NX is a multiple of 4 the constant-1 input slot would be alone in an
  // extra layer-1 k-step (and, for a multiple of 8, alone in an extra M tile of the W1 gradient): d = 8 spends 12 instead
  // of 8 DMMAs per warp and stage on layer 1 and 16 instead of 8 on the W1 gradient for it.  With BX the layer-1
  // accumulators start at b1 and the b1 gradient is the column sum of the z-bar fragments the W1-gradient MMAs load anyway.
  static constexpr bool BX = (BX_ & 1) != 0;
  // RA (bit 1 of BX_): the block's partial row [n_pm + 1] aliases the per-warp stage rings instead of owning shared memory
  // of its own -- it is only written by the flush, after the last task (C5: 17.5 of 86 KB, the difference between 2 and 3
  // blocks per SM)
  static constexpr bool RA = (BX_ & 2) != 0;
  static constexpr int D = D_, NX = NX_, HT = HT_;
  static constexpr int HW = HW_;                         // warps that share one 8-segment task, each owning HT hidden tiles
  static constexpr int HTT = HT_ * HW_;                  // hidden tiles of the whole network (8 units each)
  static constexpr int NXA = NX > 0 ? NX : 1;
  static constexpr int KS = (D + NX + (BX ? 0 : 1) + 3) / 4;   // input k-steps (state, covariates, constant-1 bias slot unless BX)
  static constexpr int NSLOT = 4 * KS;
  static constexpr int KO = (D + 3) / 4;                 // output k-steps (h-bar = W2^T k-bar)
  static constexpr int MT = (NSLOT + 7) / 8;             // M tiles of the W1 gradient (slots)
  // [segment][hidden] tiles: dense rows of 8*HT doubles with the column rotated by 4*segment when 8*HT is a power of
  // two (conflict-free for the owner's stores and for the transposed B-fragment loads, and 11% less stash traffic than
  // padding the row), else rows padded by 4.
  static constexpr bool HPOW2 = (HT & (HT - 1)) == 0;
  static constexpr int HS = HPOW2 ? 8 * HT : 8 * HT + 4;
  static constexpr int XS = MT * 8;                      // row stride of the [segment][slot] tile (dense)
  __host__ __device__ static constexpr int hpos(int seg, int col) {
    return HPOW2 ? seg * HS + ((col + 4 * seg) & (HS - 1)) : seg * HS + col;
  }
  static constexpr int REC = 8 * HS + 8 * XS;            // doubles per stage record (h tile + x tile), multiple of 2
  static constexpr int NF1 = HTT * KS, NF2 = HTT * 2, NF3 = HTT * KO, NF4 = HTT * 2;
  static constexpr int NFRAG = NF1 + NF2 + NF3 + NF4;
  static constexpr int KSTR = 12;                        // row stride of the [segment][output] k-bar tile
  static constexpr int SCRATCH = 8 * HS + 8 * KSTR;      // z-bar tile + k-bar tile
  static_assert(D <= 8, "state dimension above 8 needs a second output tile");
  __host__ __device__ static constexpr int outmap(int n) { return (n >> 1) + 4 * (n & 1); }   // N column -> output / slot
};

// per-warp shared memory (doubles): stage-record ring (or the h/x tiles for derivative matching) + z-bar/k-bar scratch
template <class MG, int LK, bool GRAD>
struct MmaSmem {
  static constexpr int WARP = GRAD ? ((LK != LK_DM ? 2 * MG::REC : MG::REC) + MG::SCRATCH) : 0;
  static constexpr int XCH = MG::HW > 1 ? 2 * MG::HW * 64 : 0;     // double-buffered exchange of the warps' partial y / x-bar
  static constexpr int BLOCK = MG::NFRAG * 32 + XCH;
};

// sum of the HW cooperating warps' partial C fragments (fixed order => identical in every warp, bitwise reproducible).
// Two alternating buffers let a warp run one exchange ahead of the slowest one with a single barrier per exchange.
template <int HW>
__device__ __forceinline__ void xchg_sum(double (&v)[2], double* s_xch, int& parity, int wl, int lane) {
  if constexpr (HW > 1) {
    double* buf = s_xch + (parity & 1) * HW * 64;
    parity ^= 1;
    *reinterpret_cast<double2*>(buf + wl * 64 + lane * 2) = make_double2(v[0], v[1]);
    __syncthreads();
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int w = 0; w < HW; ++w) {
      const double2 t = *reinterpret_cast<const double2*>(buf + w * 64 + lane * 2);
      a0 += t.x; a1 += t.y;
    }
    v[0] = a0; v[1] = a1;
  }
}

// weight of W1 for (hidden j, slot s): inputs, then the bias column, zero padding
template <class MG>
__device__ __forceinline__ double w1_slot(const DevProblem& P, const double* __restrict__ pm, int j, int s) {
  if (j >= P.H) return 0.0;
  if (s < MG::D + MG::NX) return pm[P.off_w1 + j + (long long)P.H * s];
  if (s == MG::D + MG::NX) return pm[P.off_b1 + j];
  return 0.0;
}
template <class MG>
__device__ __forceinline__ double w2_at(const DevProblem& P, const double* __restrict__ pm, int o, int j) {
  if (j >= P.H || o >= MG::D) return 0.0;
  return pm[P.off_w2 + o + (long long)P.nn_out * j];
}

// B fragments of the four products, lane-major (fragment f, lane l at [f*32 + l])
template <class MG>
__device__ __forceinline__ void build_frags(const DevProblem& P, const double* __restrict__ pm, double* __restrict__ s_w) {
  for (int idx = threadIdx.x; idx < MG::NFRAG * 32; idx += blockDim.x) {
    const int f = idx >> 5, l = idx & 31, k = l & 3, n = l >> 2;
    double v;
    if (f < MG::NF1) {                       // layer 1: B[k][n] = W1[8t+n][slot k+4e]
      const int t = f / MG::KS, e = f % MG::KS;
      v = w1_slot<MG>(P, pm, 8 * t + n, k + 4 * e);
    } else if (f < MG::NF1 + MG::NF2) {      // layer 2: B[k][n] = W2[out(n)][8t+2k+j]
      const int g = f - MG::NF1, t = g >> 1, j = g & 1;
      v = w2_at<MG>(P, pm, MG::outmap(n), 8 * t + 2 * k + j);
    } else if (f < MG::NF1 + MG::NF2 + MG::NF3) {   // h-bar: B[k][n] = W2[k+4e][8t+n]
      const int g = f - MG::NF1 - MG::NF2, t = g / MG::KO, e = g % MG::KO;
      v = w2_at<MG>(P, pm, k + 4 * e, 8 * t + n);
    } else {                                 // x-bar: B[k][n] = W1[8t+2k+j][slot out(n)], state slots only
      const int g = f - MG::NF1 - MG::NF2 - MG::NF3, t = g >> 1, j = g & 1;
      const int s = MG::outmap(n);
      v = (s < MG::D) ? w1_slot<MG>(P, pm, 8 * t + 2 * k + j, s) : 0.0;
    }
    s_w[idx] = v;
  }
}

// ---- one RHS evaluation for the warp's 8 segments; xs = this lane's input slots --------------------------------
template <class MG>
__device__ __forceinline__ void mma_stage_forward(const double* __restrict__ s_w, const double* s_tab, int lane, int wl,
                                                  double* s_xch, int& xpar,
                                                  const double (&xs)[MG::KS], const double (&b2own)[2],
                                                  double (&h)[MG::HT][2], double (&y)[2], const double* s_b1) {
  const int tb = wl * MG::HT;                 // first hidden tile of this warp
  double z[MG::HT][2];
#pragma unroll
  for (int t = 0; t < MG::HT; ++t) {
    if constexpr (MG::BX) {                   // lane (r, c) accumulates hidden units 8 (tb + t) + 2c, + 1
      const double2 b = *reinterpret_cast<const double2*>(s_b1 + 8 * (tb + t) + 2 * (lane & 3));
      z[t][0] = b.x; z[t][1] = b.y;
    } else { z[t][0] = 0.0; z[t][1] = 0.0; }
  }
#pragma unroll
  for (int e = 0; e < MG::KS; ++e) {
#pragma unroll
    for (int t = 0; t < MG::HT; ++t) dmma(z[t], xs[e], s_w[((tb + t) * MG::KS + e) * 32 + lane]);
  }
  // tanh on the 2*HT hidden units this lane owns, TB chains at a time (8 hides the 8-cycle dependent-issue latency
  // completely but needs ~40 more live registers)
  if constexpr (MG::TB == 8 && MG::HT == 4) {
    double zi[8] = {z[0][0], z[0][1], z[1][0], z[1][1], z[2][0], z[2][1], z[3][0], z[3][1]}, ho[8];
    ude_tanh_vec<8>(zi, ho, s_tab);
#pragma unroll
    for (int t = 0; t < 4; ++t) { h[t][0] = ho[2 * t]; h[t][1] = ho[2 * t + 1]; }
  } else
#pragma unroll
  for (int t = 0; t < MG::HT; t += 2) {
    if (t + 1 < MG::HT) {
      double zi[4] = {z[t][0], z[t][1], z[t + 1][0], z[t + 1][1]}, ho[4];
      ude_tanh_vec<4>(zi, ho, s_tab);
      h[t][0] = ho[0]; h[t][1] = ho[1]; h[t + 1][0] = ho[2]; h[t + 1][1] = ho[3];
    } else {
      double zi[2] = {z[t][0], z[t][1]}, ho[2];
      ude_tanh_vec<2>(zi, ho, s_tab);
      h[t][0] = ho[0]; h[t][1] = ho[1];
    }
  }
  // one accumulator per hidden tile: HT independent DMMA chains of length 2 instead of 2 chains of length HT
  double yt[MG::HT][2];
#pragma unroll
  for (int t = 0; t < MG::HT; ++t) {
    yt[t][0] = 0.0; yt[t][1] = 0.0;
    dmma(yt[t], h[t][0], s_w[(MG::NF1 + 2 * (tb + t)) * 32 + lane]);
  }
#pragma unroll
  for (int t = 0; t < MG::HT; ++t) dmma(yt[t], h[t][1], s_w[(MG::NF1 + 2 * (tb + t) + 1) * 32 + lane]);
  y[0] = 0.0; y[1] = 0.0;
#pragma unroll
  for (int t = 0; t < MG::HT; ++t) { y[0] += yt[t][0]; y[1] += yt[t][1]; }
  xchg_sum<MG::HW>(y, s_xch, xpar, wl, lane);
  y[0] += b2own[0];
  y[1] += b2own[1];
}

template <int D_, int NX_, int HT_, int HW_, int SOLVER, int LK, bool GRAD, int MINB, int TB_ = 4, int BX_ = 0>
__global__ void __launch_bounds__(kBlockThreads, MINB) ude_mma_kernel(const DevProblem P) {
  using MG = MmaGeo<D_, NX_, HT_, HW_, TB_, BX_>;
  constexpr int HW = HW_;
  static_assert(HW == 1 || HW == kBlockThreads / 32, "a task is shared by one warp or by the whole block");
  using T = Tab<SOLVER>;
  constexpr int D = D_, NX = NX_, HT = HT_, KS = MG::KS;
  __shared__ double s_tab[UDE_TANH_TAB_N];
  __shared__ __align__(8) uint64_t s_bar[kBlockThreads / 32][2];
  __shared__ __align__(16) double s_b1[MG::BX ? 8 * MG::HTT : 2];
  extern __shared__ __align__(16) double s_dyn[];
  // [n_pm + 1 (block partial) padded even][weight fragments][per warp: 2 stage records | z-bar/k-bar scratch]
  constexpr bool RA = MG::RA && GRAD;          // forward-only kernels have no per-warp region to alias
  double* s_w = RA ? s_dyn : s_dyn + (P.n_pm + 1 + 1) / 2 * 2;
  constexpr int kPerWarp = MmaSmem<MG, LK, GRAD>::WARP;
  double* s_xch = s_w + MG::NFRAG * 32;
  double* s_red = RA ? s_xch + MmaSmem<MG, LK, GRAD>::XCH : s_dyn;
  double* s_warp = s_xch + MmaSmem<MG, LK, GRAD>::XCH + (size_t)(threadIdx.x >> 5) * kPerWarp;
  double* s_stage = s_warp;
  // ring first (LK_DM: nothing), then the scratch tiles; LK_DM puts its h/x tiles behind the k-bar tile
  double* s_zs = s_warp + ((GRAD && LK != LK_DM) ? 2 * MG::REC : 0);
  double* s_ks = s_zs + 8 * MG::HS;
  uint64_t* bar = s_bar[threadIdx.x >> 5];

  const int lane = threadIdx.x & 31;
  const int r = lane >> 2, c = lane & 3;
  const long long warps_per_block = blockDim.x >> 5;
  const long long warp_global = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  // a task (8 segments) belongs to a group of HW warps; wl = this warp's slice of the hidden layer
  const int wl = (threadIdx.x >> 5) % HW;
  const long long group_global = warp_global / HW;
  const long long total_groups = (long long)gridDim.x * warps_per_block / HW;
  const bool lead = wl == 0;                  // the group's lead warp owns the loss, the uhat gradient and gb2
  int xpar = 0;
  const int S = P.substeps;
  constexpr uint32_t kRecBytes = MG::REC * 8;
  // stash: one record per RK stage
  double* __restrict__ stash = GRAD ? reinterpret_cast<double*>(P.stash) + warp_global * P.stash_rows_per_warp * 32 : nullptr;

  ude_load_tanh_tab(s_tab);
  if constexpr (!RA) for (long long k = threadIdx.x; k <= P.n_pm; k += blockDim.x) s_red[k] = 0.0;
  const double* __restrict__ pm = P.theta + (long long)D * P.model_col0[1];        // single model: [uhat | pm]
  build_frags<MG>(P, pm, s_w);
  if constexpr (MG::BX)
    for (int j = threadIdx.x; j < 8 * MG::HTT; j += blockDim.x) s_b1[j] = j < P.H ? pm[P.off_b1 + j] : 0.0;
  if (GRAD && LK != LK_DM && lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  uint32_t bar_phase = 0;
  __syncthreads();

  // which of my slots are state components; my output biases
  bool is_state[KS];
  int slot[KS];
  double b2own[2];
#pragma unroll
  for (int e = 0; e < KS; ++e) { slot[e] = c + 4 * e; is_state[e] = slot[e] < D; }
#pragma unroll
  for (int j = 0; j < 2; ++j) b2own[j] = (c + 4 * j < D) ? pm[P.off_b2 + c + 4 * j] : 0.0;

  // gradient accumulators (C fragments) -- summed over the warp's 8 segments by the MMA itself
  double GW1[MG::MT][HT][2], GW2[HT][2], gb2[2] = {0.0, 0.0};
  double gb1[HT];                              // BX: column sums of the z-bar B fragments this lane loads (tile column r)
#pragma unroll
  for (int t = 0; t < HT; ++t) {
    GW2[t][0] = GW2[t][1] = 0.0; gb1[t] = 0.0;
#pragma unroll
    for (int m = 0; m < MG::MT; ++m) GW1[m][t][0] = GW1[m][t][1] = 0.0;
  }
  double loss = 0.0;

  const double* __restrict__ uh = P.uhat_x ? P.uhat_x : P.theta;        // single model: uhat is indexed by global column
  double* __restrict__ guh = P.guhat_x ? P.guhat_x : P.grad;

  for (long long task = P.task_begin + group_global; task < P.n_tasks; task += total_groups) {
    const Seg sg = P.segs[task * 8 + r];
    const bool is_null = (sg.flags & SEG_NULL) != 0;
    const int series = sg.series;
    CovCache<NX> cur;
    cov_reset<NX>(cur);

    // assemble this lane's input slots from its state slots, the covariates and the constant
    auto make_xs = [&](const double (&U)[KS], const double (&X)[MG::NXA], double (&xs)[KS]) {
#pragma unroll
      for (int e = 0; e < KS; ++e) {
        const int s = slot[e];
        double v = 0.0;
        if (s < D) v = U[e];
        else if (s == D + NX) v = 1.0;
#pragma unroll
        for (int q = 0; q < NX; ++q) if (s == D + q) v = X[q];
        xs[e] = v;
      }
    };

    if (LK == LK_DM) {
      const long long col = sg.col;
      double U[KS], X[MG::NXA], xs[KS], h[HT][2], y[2];
      X[0] = 0.0;
#pragma unroll
      for (int e = 0; e < KS; ++e) U[e] = is_state[e] ? P.dm_u[D * col + slot[e]] : 0.0;
      const double ts = P.times[col];
      if (NX > 0) cov_eval<NX>(P, series, ts, cur, X);
      make_xs(U, X, xs);
      mma_stage_forward<MG>(s_w, s_tab, lane, wl, s_xch, xpar, xs, b2own, h, y, s_b1);
      const double w = is_null ? 0.0 : 1.0;
      double kb[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool st = (c + 4 * j) < D;
        const double res = st ? P.dm_dudt[D * col + c + 4 * j] - y[j] : 0.0;
        loss = fma(w * res, res, loss);
        kb[j] = -2.0 * w * res;
      }
      if (GRAD) {
        // ---- VJP of one stage (same code as the reverse sweep below, without a stage record) -------------------
        double hb[HT][2], zb[HT][2];
#pragma unroll
        for (int t = 0; t < HT; ++t) { hb[t][0] = hb[t][1] = 0.0; }
#pragma unroll
        for (int e = 0; e < MG::KO; ++e) {
#pragma unroll
          for (int t = 0; t < HT; ++t) dmma(hb[t], kb[e], s_w[(MG::NF1 + MG::NF2 + (wl * HT + t) * MG::KO + e) * 32 + lane]);
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < HT; ++t) {
          zb[t][0] = hb[t][0] * fma(-h[t][0], h[t][0], 1.0);
          zb[t][1] = hb[t][1] * fma(-h[t][1], h[t][1], 1.0);
          s_zs[MG::hpos(r, 8 * t + c)] = zb[t][0]; s_zs[MG::hpos(r, 8 * t + c + 4)] = zb[t][1];
        }
        s_ks[r * MG::KSTR + c] = kb[0]; s_ks[r * MG::KSTR + 4 + c] = kb[1];
        if (lead) { gb2[0] += kb[0]; gb2[1] += kb[1]; }
        // no stage record in this loss family: the h and x tiles are staged behind the k-bar tile
        double* s_hs = s_ks + 8 * MG::KSTR;
        double* s_xs = s_hs + 8 * MG::HS;
#pragma unroll
        for (int t = 0; t < HT; ++t) { s_hs[MG::hpos(r, 8 * t + c)] = h[t][0]; s_hs[MG::hpos(r, 8 * t + c + 4)] = h[t][1]; }
#pragma unroll
        for (int e = 0; e < KS; ++e) s_xs[r * MG::XS + slot[e]] = xs[e];
        __syncwarp();
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const int seg = 4 * s2 + c;
          double aX[MG::MT];
#pragma unroll
          for (int m = 0; m < MG::MT; ++m) aX[m] = s_xs[seg * MG::XS + r + 8 * m];
          const double aK = s_ks[seg * MG::KSTR + r];
#pragma unroll
          for (int t = 0; t < HT; ++t) {
            const double bZ = s_zs[MG::hpos(seg, 8 * t + r)];
            const double bH = s_hs[MG::hpos(seg, 8 * t + r)];
#pragma unroll
            for (int m = 0; m < MG::MT; ++m) dmma(GW1[m][t], aX[m], bZ);
            dmma(GW2[t], aK, bH);
            if constexpr (MG::BX) gb1[t] += bZ;
          }
        }
        __syncwarp();
      }
    } else {
      // ---- a segment = [pre-roll step] + maxL intervals of S steps (same structure as ude_seg_kernel) ---------------
      const bool multi = P.loss_kind == UDE_LOSS_MULTI_SHOOT;
      const bool skip = is_null || (LK == LK_STATE && (sg.flags & SEG_SKIP_PROC));
      long long c0;
      int L, maxL, pre, first_scored;
      bool from_theta, has_prev = true;
      double wpt = 0.0, eps = 0.0;
      if (LK == LK_STATE) {
        const long long col = sg.col;
        has_prev = !is_null && !(sg.flags & SEG_NO_PREV);
        c0 = has_prev ? col - 1 : col;
        L = (skip || !has_prev) ? 0 : 1; maxL = 1; pre = 0; first_scored = 0; from_theta = true;
      } else {
        c0 = sg.col;
        L = is_null ? 0 : sg.L;
        maxL = (int)__reduce_max_sync(0xffffffffu, (unsigned)L);
        from_theta = multi || P.shoot_from_theta;
        first_scored = multi ? 0 : P.shoot_first_scored;
        wpt = is_null ? 0.0 : (multi ? 1.0 / ((double)D * (double)P.lengths[series]) : P.shoot_weight[0]);
        eps = multi ? P.preroll_eps : 0.0;
        pre = eps > 0.0 ? 1 : 0;
      }
      const int nsteps = pre + maxL * S;
      // the interval loop reads times and this lane's data entries once per interval, each a cold HBM access on the
      // critical path: start them all now
      for (int i = lane & 3; i <= L + 1; i += 4) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.times + c0 + min(i, L)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.data + (long long)D * (c0 + min(i, L))));
      }
      double u[KS];
#pragma unroll
      for (int e = 0; e < KS; ++e) u[e] = is_state[e] ? (from_theta ? uh[D * c0 + slot[e]] : P.data[D * c0 + slot[e]]) : 0.0;

      // ------------------------------------------ forward sweep ----------------------------------------------------
      {
        int i = pre ? -1 : 0, n = pre ? S - 1 : 0;
        double t0 = P.times[c0] - eps, hs = is_null ? 0.0 : eps, nd = 0.0;
        for (int k = 0; k < nsteps; ++k) {
          if (k >= pre && n == 0) {
            const int ic = min(i, L);
            if (LK == LK_TRAJ) {
              const double w = (i <= L && i >= first_scored) ? wpt : 0.0;
#pragma unroll
              for (int e = 0; e < KS; ++e) if (is_state[e]) {
                const double res = P.data[D * (c0 + ic) + slot[e]] - u[e];
                loss = fma(w * res, res, loss);
                if (!GRAD && lead && P.traj_out != nullptr && i <= L && !is_null) P.traj_out[D * (c0 + i) + slot[e]] = u[e];
              }
            }
            t0 = P.times[c0 + ic];
            hs = (i < L) ? (P.times[c0 + ic + 1] - t0) / S : 0.0;
            nd = 0.0;
          }
          const double t = fma(nd, hs, t0);
          // one explicit RK step on this lane's state slots
          double kst[T::S][KS], acc[KS];
#pragma unroll
          for (int e = 0; e < KS; ++e) acc[e] = 0.0;
#pragma unroll
          for (int st = 0; st < T::S; ++st) {
            double U[KS], X[MG::NXA], xs[KS], h[HT][2], y[2];
            X[0] = 0.0;
#pragma unroll
            for (int e = 0; e < KS; ++e) {
              double a = 0.0;
#pragma unroll
              for (int l = 0; l < st; ++l) if (T::a(st, l) != 0.0) a = fma(T::a(st, l), kst[l][e], a);
              U[e] = (st == 0) ? u[e] : fma(hs, a, u[e]);
            }
            const double ts = t + T::c(st) * hs;
            if (NX > 0) cov_eval<NX>(P, series, ts, cur, X);
            make_xs(U, X, xs);
            mma_stage_forward<MG>(s_w, s_tab, lane, wl, s_xch, xpar, xs, b2own, h, y, s_b1);
            if (GRAD) {
              // stage record -> shared memory -> TMA bulk store to the stash (two-buffer ring, one record per stage)
              const int rec = k * T::S + st;
              double* sbuf = s_stage + (rec & 1) * MG::REC;
              if (lane == 0) bulk_wait_read<1>();
              __syncwarp();
#pragma unroll
              for (int t2 = 0; t2 < HT; ++t2) { sbuf[MG::hpos(r, 8 * t2 + c)] = h[t2][0]; sbuf[MG::hpos(r, 8 * t2 + c + 4)] = h[t2][1]; }
#pragma unroll
              for (int e = 0; e < KS; ++e) sbuf[8 * MG::HS + r * MG::XS + slot[e]] = xs[e];
              fence_async_smem();
              __syncwarp();
              if (lane == 0) { bulk_store(stash + (long long)rec * MG::REC, sbuf, kRecBytes); bulk_commit(); }
            }
#pragma unroll
            for (int e = 0; e < KS; ++e) {
              const double du = (is_state[e] && e < 2) ? y[e < 2 ? e : 0] : 0.0;
              kst[st][e] = du; acc[e] = fma(T::b(st), du, acc[e]);
            }
          }
#pragma unroll
          for (int e = 0; e < KS; ++e) u[e] = fma(hs, acc[e], u[e]);
          nd += 1.0;
          if (++n == S) { n = 0; ++i; }
        }
        if (GRAD) {
          if (lane == 0) {
            bulk_wait_all<0>();
            if (nsteps > 0) {
              const int rec = nsteps * T::S - 1, b = rec & 1;
              mbar_expect_tx(&bar[b], kRecBytes);
              bulk_load(s_stage + b * MG::REC, stash + (long long)rec * MG::REC, kRecBytes, &bar[b]);
            }
          }
          __syncwarp();
        }
      }
      // ------------------------------------------ loss at the last point ----------------------------------------------
      double lam[KS], gc[KS], gp[KS];
#pragma unroll
      for (int e = 0; e < KS; ++e) { lam[e] = 0.0; gc[e] = 0.0; gp[e] = 0.0; }
      if (LK == LK_STATE) {
        const long long col = sg.col;
        const double dt = P.times[col] - P.times[c0];
        const double wsc = (L == 1) ? P.proc_scale[0] * (P.proc_inv_dt ? 1.0 / dt : 1.0) : 0.0;
        const double os = is_null ? 0.0 : P.obs_scale[0];
        const bool obs_first = (sg.flags & SEG_OBS_FIRST) && has_prev;
        // residual vectors are spread over the quad: gather them (once per segment) for the d x d quadratic forms
        double rown[2], eown[2], fown[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const bool st = (c + 4 * j) < D;
          rown[j] = st ? uh[D * col + c + 4 * j] - u[j] : 0.0;
          eown[j] = st ? P.data[D * col + c + 4 * j] - uh[D * col + c + 4 * j] : 0.0;
          fown[j] = (st && obs_first) ? P.data[D * c0 + c + 4 * j] - uh[D * c0 + c + 4 * j] : 0.0;
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
        for (int e = 0; e < KS; ++e) if (is_state[e]) {
#pragma unroll
          for (int b = 0; b < D; ++b) if (b == slot[e]) {
            lam[e] = -wsc * sym[b]; gc[e] = wsc * sym[b] - os * es[b]; gp[e] = -os * fs[b];
          }
        }
        if (!GRAD && lead && P.traj_out != nullptr && !is_null) {
#pragma unroll
          for (int e = 0; e < KS; ++e) if (is_state[e]) {
            if (L == 1) P.traj_out[D * col + slot[e]] = u[e];
            if (!has_prev || (sg.flags & SEG_OBS_FIRST)) P.traj_out[D * c0 + slot[e]] = uh[D * c0 + slot[e]];
          }
        }
      } else {
        const double w = (maxL <= L && maxL >= first_scored) ? wpt : 0.0;
        const int ic = min(maxL, L);
#pragma unroll
        for (int e = 0; e < KS; ++e) if (is_state[e]) {
          const double res = P.data[D * (c0 + ic) + slot[e]] - u[e];
          loss = fma(w * res, res, loss);
          lam[e] = -2.0 * w * res;
          if (!GRAD && lead && P.traj_out != nullptr && maxL <= L && !is_null && !(multi && L == P.pred_length))
            P.traj_out[D * (c0 + L) + slot[e]] = u[e];
        }
      }
      // ------------------------------------------ reverse sweep ----------------------------------------------------------
      if (GRAD) {
        int i = maxL - 1, n = S - 1;
        double t0 = 0.0, hs = 0.0;
        for (int k = nsteps - 1; k >= 0; --k) {
          if (k >= pre) {
            if (n == S - 1) {
              const int ic = min(i, L);
              t0 = P.times[c0 + ic];
              hs = (i < L) ? (P.times[c0 + ic + 1] - t0) / S : 0.0;
            }
          } else { hs = is_null ? 0.0 : eps; }
          double Ub[T::S][KS], lamacc[KS], U0[KS];
#pragma unroll
          for (int e = 0; e < KS; ++e) { lamacc[e] = 0.0; U0[e] = 0.0; }
#pragma unroll
          for (int ii = 0; ii < T::S; ++ii) {
            const int st = T::S - 1 - ii;
            const int rec = k * T::S + st, b = rec & 1;
            if (rec > 0 && lane == 0) {          // prefetch the next record to be reversed into the other buffer
              mbar_expect_tx(&bar[b ^ 1], kRecBytes);
              bulk_load(s_stage + (b ^ 1) * MG::REC, stash + (long long)(rec - 1) * MG::REC, kRecBytes, &bar[b ^ 1]);
            }
            mbar_wait(&bar[b], (bar_phase >> b) & 1u);
            bar_phase ^= (1u << b);
            const double* __restrict__ sbuf = s_stage + b * MG::REC;
            double h[HT][2], xs[KS];
#pragma unroll
            for (int t = 0; t < HT; ++t) { h[t][0] = sbuf[MG::hpos(r, 8 * t + c)]; h[t][1] = sbuf[MG::hpos(r, 8 * t + c + 4)]; }
#pragma unroll
            for (int e = 0; e < KS; ++e) xs[e] = sbuf[8 * MG::HS + r * MG::XS + slot[e]];
            // k-bar on my output slots (outputs o = c + 4j coincide with state slots e = j)
            double kb[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              double a = 0.0;
              if (j < KS) {
                a = T::b(st) * lam[j < KS ? j : 0];
#pragma unroll
                for (int l = st + 1; l < T::S; ++l) if (T::a(l, st) != 0.0) a = fma(T::a(l, st), Ub[l][j < KS ? j : 0], a);
              }
              kb[j] = ((c + 4 * j) < D) ? hs * a : 0.0;
            }
            double hb[HT][2], zb[HT][2];
#pragma unroll
            for (int t = 0; t < HT; ++t) { hb[t][0] = hb[t][1] = 0.0; }
#pragma unroll
            for (int e = 0; e < MG::KO; ++e) {
#pragma unroll
              for (int t = 0; t < HT; ++t) dmma(hb[t], kb[e], s_w[(MG::NF1 + MG::NF2 + (wl * HT + t) * MG::KO + e) * 32 + lane]);
            }
#pragma unroll
            for (int t = 0; t < HT; ++t) {
              zb[t][0] = hb[t][0] * fma(-h[t][0], h[t][0], 1.0);
              zb[t][1] = hb[t][1] * fma(-h[t][1], h[t][1], 1.0);
              s_zs[MG::hpos(r, 8 * t + c)] = zb[t][0]; s_zs[MG::hpos(r, 8 * t + c + 4)] = zb[t][1];
            }
            s_ks[r * MG::KSTR + c] = kb[0]; s_ks[r * MG::KSTR + 4 + c] = kb[1];
            if (lead) { gb2[0] += kb[0]; gb2[1] += kb[1]; }
            // x-bar = W1^T z-bar on the state slots
            double xt[HT][2];
#pragma unroll
            for (int t = 0; t < HT; ++t) {
              xt[t][0] = 0.0; xt[t][1] = 0.0;
              dmma(xt[t], zb[t][0], s_w[(MG::NF1 + MG::NF2 + MG::NF3 + 2 * (wl * HT + t)) * 32 + lane]);
            }
#pragma unroll
            for (int t = 0; t < HT; ++t) dmma(xt[t], zb[t][1], s_w[(MG::NF1 + MG::NF2 + MG::NF3 + 2 * (wl * HT + t) + 1) * 32 + lane]);
            double xsum[2] = {0.0, 0.0};
#pragma unroll
            for (int t = 0; t < HT; ++t) { xsum[0] += xt[t][0]; xsum[1] += xt[t][1]; }
            xchg_sum<HW>(xsum, s_xch, xpar, wl, lane);
#pragma unroll
            for (int e = 0; e < KS; ++e) {
              const double ub = (e < 2 && is_state[e]) ? xsum[e < 2 ? e : 0] : 0.0;
              Ub[st][e] = ub; lamacc[e] += ub;
              if (st == 0) U0[e] = xs[e];
            }
            __syncwarp();                       // z-bar / k-bar tiles complete
            // weight gradients: contraction over the warp's 8 segments
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
              const int seg = 4 * s2 + c;
              double aX[MG::MT];
#pragma unroll
              for (int m = 0; m < MG::MT; ++m) aX[m] = sbuf[8 * MG::HS + seg * MG::XS + r + 8 * m];
              const double aK = s_ks[seg * MG::KSTR + r];
#pragma unroll
              for (int t = 0; t < HT; ++t) {
                const double bZ = s_zs[MG::hpos(seg, 8 * t + r)];
                const double bH = sbuf[MG::hpos(seg, 8 * t + r)];
#pragma unroll
                for (int m = 0; m < MG::MT; ++m) dmma(GW1[m][t], aX[m], bZ);
                dmma(GW2[t], aK, bH);
                if constexpr (MG::BX) gb1[t] += bZ;
              }
            }
            __syncwarp();                       // everyone is done with buffer b and the scratch tiles
            // the record is dead: drop its (dirty) L2 lines instead of letting them be written back to HBM
            if (P.stash_discard) {
              for (int ln = lane; ln < (int)(kRecBytes / 128); ln += 32)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(stash + (long long)rec * MG::REC + ln * 16) : "memory");
            }
          }
#pragma unroll
          for (int e = 0; e < KS; ++e) lam[e] += lamacc[e];
          if (k >= pre) {
            if (n == 0) {
              if (LK == LK_TRAJ) {
                const int ic = min(i, L);
                const double w = (i <= L && i >= first_scored) ? wpt : 0.0;
#pragma unroll
                for (int e = 0; e < KS; ++e) if (is_state[e]) lam[e] = fma(-2.0 * w, P.data[D * (c0 + ic) + slot[e]] - U0[e], lam[e]);
              }
              --i; n = S - 1;
            } else --n;
          }
        }
        if (!is_null && lead) {
#pragma unroll
          for (int e = 0; e < KS; ++e) if (is_state[e]) {
            if (LK == LK_STATE) {
              atomicAdd(guh + D * sg.col + slot[e], gc[e]);
              if (has_prev) atomicAdd(guh + D * c0 + slot[e], gp[e] + lam[e]);
            } else if (from_theta) { if (P.guhat_x) guh[D * c0 + slot[e]] = lam[e]; else atomicAdd(guh + D * c0 + slot[e], lam[e]); }
          }
        }
      }
    }
  }

  // ---- flush: the C fragments already hold sums over the warp's segments; warps add into the block row in order ----
  {
    // gb2 and the loss still need the sum over the 8 segments (lane bits 2..4), the loss also over the quad
    double l = lead ? loss : 0.0;          // the cooperating warps computed identical losses: count the lead's
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) { gb2[0] += __shfl_xor_sync(0xffffffffu, gb2[0], off); gb2[1] += __shfl_xor_sync(0xffffffffu, gb2[1], off); }
    if constexpr (MG::BX && GRAD) {        // lanes (r, 0..3) loaded the 8 segments of tile column r between them
#pragma unroll
      for (int t = 0; t < HT; ++t) { gb1[t] += __shfl_xor_sync(0xffffffffu, gb1[t], 1); gb1[t] += __shfl_xor_sync(0xffffffffu, gb1[t], 2); }
    }
    const int nw = blockDim.x >> 5, wib = threadIdx.x >> 5;
    if constexpr (RA) {                     // every warp is done with its rings (bulk stores drained, loads consumed)
      __syncthreads();
      for (long long k = threadIdx.x; k <= P.n_pm; k += blockDim.x) s_red[k] = 0.0;
      __syncthreads();
    }
    for (int w = 0; w < nw; ++w) {
      if (wib == w) {
        if (GRAD) {
#pragma unroll
          for (int t = 0; t < HT; ++t) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int ncol = 2 * c + j;              // tile column n holds the hidden unit of lane (n & 3), value (n >> 2)
              const int hid = 8 * (wl * HT + t) + 2 * (ncol & 3) + (ncol >> 2);
              if (hid < P.H) {
#pragma unroll
                for (int m = 0; m < MG::MT; ++m) {
                  const int s = r + 8 * m;            // slot: input column, or the bias
                  if (s < D + NX) s_red[P.off_w1 + hid + (long long)P.H * s] += GW1[m][t][j];
                  else if (!MG::BX && s == D + NX) s_red[P.off_b1 + hid] += GW1[m][t][j];   // BX: rows >= 4 KS of the x tile are never written
                }
                if (r < D) s_red[P.off_w2 + r + (long long)P.nn_out * hid] += GW2[t][j];
              }
            }
          }
          if (r == 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j) if (c + 4 * j < D) s_red[P.off_b2 + c + 4 * j] += gb2[j];
          }
          if constexpr (MG::BX) {
            if (c == 0) {
#pragma unroll
              for (int t = 0; t < HT; ++t) {         // tile column r holds hidden unit 2 (r & 3) + (r >> 2) of the tile
                const int hid = 8 * (wl * HT + t) + 2 * (r & 3) + (r >> 2);
                if (hid < P.H) s_red[P.off_b1 + hid] += gb1[t];
              }
            }
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
