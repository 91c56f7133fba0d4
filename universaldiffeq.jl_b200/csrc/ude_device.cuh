// Device-side problem description, segment table entry, RK tableaux and sub-warp group collectives.
//
// Layout in HBM (DESIGN.md "Data layout"):
//   times[n_cols], data[d x n_cols] column-major, theta = concat_m [uhat_m | pm_m], grad same shape,
//   covariate knots CSR (knot_off, knot_t, knot_x, knot_inv = 1/(t[i+1]-t[i])),
//   segs[n_tasks * SPW]: 16-byte work items sorted by (model, series, time), padded per model so that one
//   warp-task (SPW = 32/G consecutive segments) never mixes models,
//   stash[total_warps][rows][32]: per-warp scratch for the reverse sweep, one 256-byte row = one double per lane.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ude {

constexpr int kMaxD = 8;
constexpr int kBlockThreads = 128;

enum : uint16_t { SEG_NULL = 1, SEG_OBS_FIRST = 2, SEG_SKIP_PROC = 4,
                  SEG_NO_PREV = 8 };   // state-space segment of a one-point series: no interval ends at its column

struct Seg {
  int32_t col;      // STATE_SPACE: END column of the interval; otherwise first column of the block
  int32_t series;   // global series index
  int16_t L;        // intervals in the block (trajectory kinds)
  uint16_t flags;
  int32_t model;
};
static_assert(sizeof(Seg) == 16, "Seg must stay 16 bytes");

struct DevProblem {
  // model
  int d, n_cov, nn_in, H, nn_out, n_scalars, has_series_param, substeps;
  long long off_w1, off_b1, off_w2, off_b2, off_scalar[8], off_series_param, n_pm;
  double time_scale, time_shift;
  // loss
  int loss_kind, pred_length, proc_inv_dt, shoot_from_theta, shoot_first_scored, remove_ends;
  double preroll_eps;
  double A[kMaxD * kMaxD], Asym[kMaxD * kMaxD], B[kMaxD * kMaxD], Bsym[kMaxD * kMaxD];
  // data (device pointers)
  const long long* starts; const long long* lengths;
  const long long* theta_base; const long long* model_col0; const long long* model_series0;
  const double* times; const double* data;
  const double* obs_scale; const double* proc_scale; const double* shoot_weight;
  const long long* knot_off; const double* knot_t; const double* knot_x; const double* knot_inv;
  const double* dm_u; const double* dm_dudt;
  // work
  const Seg* segs; long long task_begin, n_tasks; int n_models;   // this launch covers tasks [task_begin, n_tasks)
  long long n_cols_total;     // columns of data / uhat over all models (bounds the next-task prefetch guesses)
  // io
  const double* theta; double* grad; double* traj_out;
  // Single-model trajectory losses read only the block-start columns of uhat and each of those columns receives exactly
  // one gradient contribution (LFC.jl:281-296, :665-680).  When uhat_x / guhat_x are set, those columns are read from /
  // written (plain stores, no atomics) to these bases instead of theta / grad -- the host-buffer path points them at the
  // caller's pinned theta and gradient, so nothing but the block starts ever crosses PCIe (ude_loss_grad, zero-copy).
  const double* uhat_x; double* guhat_x;
  double* block_part;   // [grid][n_pm + 1]: per-block gradient of process_model and loss (n_models == 1)
  double* model_loss;   // [n_models], atomics                                       (n_models  > 1)
  void* stash; long long stash_rows_per_warp;      // rows of 32 elements of the kernel's arithmetic type
  int stash_discard;    // 1: records are discarded from L2 (no write-back) once the reverse sweep has consumed them
};

// ------------------------------------------------------------------------------------------------
// Butcher tableaux as constexpr functions so zero coefficients vanish at compile time.
template <int SOLVER> struct Tab;
template <> struct Tab<0> {   // classic RK4
  static constexpr int S = 4;
  __host__ __device__ static constexpr double c(int i) { constexpr double v[4] = {0.0, 0.5, 0.5, 1.0}; return v[i]; }
  __host__ __device__ static constexpr double b(int i) { constexpr double v[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0}; return v[i]; }
  __host__ __device__ static constexpr double a(int i, int j) {
    constexpr double v[4][4] = {{0, 0, 0, 0}, {0.5, 0, 0, 0}, {0, 0.5, 0, 0}, {0, 0, 1.0, 0}};
    return v[i][j];
  }
};
template <> struct Tab<1> {   // Tsit5 stages 1..6 (OrdinaryDiffEq Tsit5ConstantCache; b = row 7 of a, b7 = 0)
  static constexpr int S = 6;
  __host__ __device__ static constexpr double c(int i) {
    constexpr double v[6] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0}; return v[i];
  }
  __host__ __device__ static constexpr double b(int i) {
    constexpr double v[6] = {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774};
    return v[i];
  }
  __host__ __device__ static constexpr double a(int i, int j) {
    constexpr double v[6][6] = {
        {0, 0, 0, 0, 0, 0},
        {0.161, 0, 0, 0, 0, 0},
        {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
        {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
        {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
        {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0}};
    return v[i][j];
  }
};

template <> struct Tab<2> {   // explicit Euler (spline gradient matching: one RHS evaluation at the left knot)
  static constexpr int S = 1;
  __host__ __device__ static constexpr double c(int) { return 0.0; }
  __host__ __device__ static constexpr double b(int) { return 1.0; }
  __host__ __device__ static constexpr double a(int, int) { return 0.0; }
};

// ------------------------------------------------------------------------------------------------
// All-reduce of N doubles over the G lanes of a sub-warp group (G a power of two, groups aligned).
// Reduce-scatter (halving the payload each butterfly step) followed by an all-gather: for N = 4, G = 16
// this is 8 double shuffles instead of the 16 of a per-value butterfly.  SHFL issues at 1 warp-instruction
// per clock per SM on B200 (profiles/microbench_r01.json), i.e. one double shuffle costs as much pipe
// time as four DFMAs, so shuffle count is what this routine minimises.  Every lane of the warp must call
// it (full mask); all lanes of a group end with bit-identical sums.
__host__ __device__ constexpr int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }
__host__ __device__ constexpr int ilog2(int n) { int l = 0; while (n > 1) { n >>= 1; ++l; } return l; }

template <int G, int N, class TV>
__device__ __forceinline__ void group_allreduce(TV (&v)[N], int lane) {
  if constexpr (G == 1) { return; }
  else if constexpr (G == 2) {
    // two lanes: the plain exchange is N double shuffles and ONE shuffle latency; reduce-scatter + gather would be
    // NP/2 + N shuffles, 2 NP selects and two dependent shuffle latencies (measured on the ten-unit kernels)
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
  }
  else {
    constexpr int NP = next_pow2(N);
    TV w[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) w[i] = (i < N) ? v[i] : TV(0);
    // reduce-scatter
    int cnt = NP;
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) {
      if (cnt > 1) {
        const int half = cnt / 2;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < NP / 2; ++i) {
          if (i < half) {
            TV send = up ? w[i] : w[i + half];
            TV keep = up ? w[i + half] : w[i];
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        cnt = half;
      } else {
        w[0] += __shfl_xor_sync(0xffffffffu, w[0], off);
      }
    }
    // gather: after the reduce-scatter, element e = q * cnt_final + r (r < cnt_final) lives in slot r of every lane
    // whose split bits (lane & G/2, lane & G/4, ...) spell q, most significant split first.  Each lane fetches all N
    // sums with independent indexed shuffles: one level of latency, no selects.
    constexpr int cnt_final = (NP / G) > 1 ? (NP / G) : 1;
    constexpr int ns = ilog2(NP / cnt_final);
    constexpr int split_mask = (G - 1) & ~((G >> ns) - 1);     // lane bits consumed by the splits
    const int keep = lane & ~split_mask;
    TV out[N];
#pragma unroll
    for (int e = 0; e < N; ++e) {
      constexpr int dummy = 0; (void)dummy;
      const int qe = e / cnt_final, re = e % cnt_final;
      int src = keep;
#pragma unroll
      for (int s2 = 0; s2 < ns; ++s2) if ((qe >> (ns - 1 - s2)) & 1) src |= (G >> (s2 + 1));
      out[e] = __shfl_sync(0xffffffffu, w[re], src);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = out[i];
  }
}

// ------------------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) between a warp's shared-memory staging buffers and its stash in global
// memory, completion through an mbarrier (loads) or a bulk async-group (stores).  One lane issues, the warp consumes.
// hint only: pull the line holding p towards L1 (the next task's inputs while this task's reverse sweep runs)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE_%=;\n"
      "bra LAB_WAIT_%=;\n"
      "LAB_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace ude
