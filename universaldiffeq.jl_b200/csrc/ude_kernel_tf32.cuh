// FP32-mode tensor-core kernel for wide NODE right-hand sides (BASELINE config C5: d = 8, hidden 128): the per-stage MLP
// of SIXTEEN segments per block runs on `mma.sync.m16n8k8.tf32` with the 3xTF32 split (a = a_hi + a_lo, both TF32;
// a b ~ a_lo b_hi + a_hi b_lo + a_hi b_hi), FP32 accumulators -- FP32-level accuracy from the TF32 tensor path.  State, RK
// combinations, tanh and the stash are FP32; times, residuals, loss sums, adjoint seeds and every cross-block reduction
// stay FP64 (the convention of the optional FP32 mode; bar 1e-4 against the FP64 oracle).
//
// Mapping (m16n8k8: A 16x8 row-major a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); B 8x8 b0 (t, g) b1 (t+4, g);
// C 16x8 c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1); g = lane >> 2, t = lane & 3):
//   * a block works on one task of 16 segments; rows g and g+8 of every fragment are the lane's two segments ("halves");
//     lane (g, t) owns state / output slots {t, t+4} of both -- exactly what it must supply as A for the first layer (one
//     k-step: K = 8 = d) and for h-bar = W2^T k-bar;
//   * the hidden layer is split over the block's four warps, 32 units (four 8-wide n-tiles) each; the partial W2 h and
//     W1^T z-bar sums are exchanged through shared memory (double-buffered, one __syncthreads per exchange, fixed order);
//   * contraction indices are permuted so a lane's C fragment is its next A fragment: k-position p of a hidden k-step is
//     hidden unit 2 (p & 3) + (p >> 2) of the tile, output column n of an 8-wide N tile is output / slot (n >> 1) + 4 (n & 1);
//     b1 rides in the accumulator of the first layer, b2 is added by the owner;
//   * weight gradients contract over the 16 segments (two k-steps) with M = hidden (two m-tiles per warp), N = slots /
//     outputs (8: no padding); the transposed operands come from padded shared-memory tiles (conflict-free both ways);
//     bias gradients are per-lane sums reduced by shuffles at the flush;
//   * weights live in shared memory as pre-split (hi, lo) B fragments, one LDS.128 per (tile, product);
//   * per-warp stage records (h tile 16 x 32 + x tile) go to the stash by TMA bulk copies, as in the FP64 kernels.
// Loss families LK_STATE and LK_TRAJ; the per-interval prefetch of times / data is the one of ude_kernel_mma2.cuh.
#pragma once
#include "ude_kernel_mma2.cuh"

namespace ude {

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + lo with hi the top 10 mantissa bits (truncated) and lo the exact remainder.  The tensor core reads only the TF32
// part of an FP32 register (the low 13 mantissa bits are ignored), so lo needs no conversion: 2 instructions per split
// (LOP3 + FADD).  `cvt.rna.tf32.f32` is emulated on sm_100a with FSETP / LOP3 / IADD / SEL sequences -- the round-1 profile of
// this kernel (profiles/ncu_c5_tf32_r02c.txt) had FSETP alone at 14 % of the issued instructions.  Truncation instead of
// round-to-nearest costs one bit: 2^-21 per operand instead of 2^-22.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// FP32 tanh for this kernel: 1 - 2 / (exp(2|x|) + 1) with ex2.approx and rcp.approx (absolute error ~1e-7; the FP32-mode
// bar is 1e-4): 6 instructions instead of ude_tanhf's ~20.  exp overflow -> inf -> rcp 0 -> 1, no clamp needed.
__device__ __forceinline__ float tf_tanhf(float x) {
  float e, y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(x) * 2.8853900817779268f));      // 2 log2(e)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(e + 1.0f));
  return copysignf(fmaf(-2.0f, y, 1.0f), x);
}
// c += A B, A given in FP32 (split here), B pre-split: small terms first
__device__ __forceinline__ void mma3(float (&c)[4], const float (&a)[4], const uint4 b /* b0hi b1hi b0lo b1lo */) {
  uint32_t ah[4], al[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
  mma_tf32(c, al, b.x, b.y);
  mma_tf32(c, ah, b.z, b.w);
  mma_tf32(c, ah, b.x, b.y);
}
// both operands in FP32
__device__ __forceinline__ void mma3_ff(float (&c)[4], const float (&a)[4], float b0, float b1) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
  split_tf32(b0, bh[0], bl[0]); split_tf32(b1, bh[1], bl[1]);
  mma_tf32(c, al, bh[0], bh[1]);
  mma_tf32(c, ah, bl[0], bl[1]);
  mma_tf32(c, ah, bh[0], bh[1]);
}

__host__ __device__ constexpr int tf_kmap(int p) { return 2 * (p & 3) + (p >> 2); }      // k-position -> hidden unit of the tile
__host__ __device__ constexpr int tf_omap(int n) { return (n >> 1) + 4 * (n & 1); }      // N column -> output / slot

template <int D_, int HTW_>
struct Tf32Geo {
  static constexpr int D = D_, HTW = HTW_, HW = kBlockThreads / 32;
  static constexpr int HWARP = 8 * HTW;                 // hidden units per warp
  static constexpr int HS = HWARP + 8;                  // row stride (floats) of the [segment][hidden] tiles: conflict-free
  static constexpr int XS = 12;                         // row stride of the [segment][slot] tiles
  static constexpr int HTILE = 16 * HS, XTILE = 16 * XS;
  static constexpr int REC = HTILE + XTILE;             // floats per warp and stage (h tile + x tile), multiple of 4
  static constexpr int SCR = HTILE + XTILE;             // z-bar tile + k-bar tile
  static constexpr int NFRAG = 4 * HTW;                 // uint4 fragments per lane and warp: layer 1, layer 2, h-bar, x-bar
  static constexpr int WARP_F = (2 * REC + SCR);        // floats per warp (grad kernels)
  static constexpr int XCH_F = 2 * HW * 32 * 4;         // exchange buffers (two parities)
  static constexpr int BLOCK_BYTES = HW * NFRAG * 32 * 16 + HW * HWARP * 4 + XCH_F * 4;   // fragments + b1 + exchange
  static_assert(D == 8, "one k-step of eight inputs: d = 8");
  static_assert(REC % 4 == 0, "TMA bulk copies need 16-byte sizes");
};

// sum of the four warps' partial C fragments, fixed order (identical in every warp)
__device__ __forceinline__ void tf_xchg(float (&v)[4], float* s_xch, int& parity, int wl, int lane) {
  constexpr int HW = kBlockThreads / 32;
  float* buf = s_xch + (parity & 1) * HW * 128;
  parity ^= 1;
  *reinterpret_cast<float4*>(buf + wl * 128 + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
  __syncthreads();
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int w = 0; w < HW; ++w) {
    const float4 q = *reinterpret_cast<const float4*>(buf + w * 128 + lane * 4);
    a0 += q.x; a1 += q.y; a2 += q.z; a3 += q.w;
  }
  v[0] = a0; v[1] = a1; v[2] = a2; v[3] = a3;
}

template <int D_, int HTW_, int SOLVER, int LK, bool GRAD, int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB) ude_tf32_kernel(const DevProblem P) {
  using TG = Tf32Geo<D_, HTW_>;
  using T = Tab<SOLVER>;
  static_assert(LK == LK_STATE || LK == LK_TRAJ, "derivative matching stays with the lane-per-hidden-unit kernel");
  constexpr int D = D_, HTW = HTW_, HW = TG::HW, KSU = 2;
  __shared__ __align__(8) uint64_t s_bar[kBlockThreads / 32][2];
  extern __shared__ __align__(16) double s_dyn[];
  // [n_pm + 1 doubles (block partial), padded even][fragments uint4][b1 floats][exchange floats][per warp: 2 records | scratch]
  double* s_red = s_dyn;
  uint4* s_frag = reinterpret_cast<uint4*>(s_dyn + (P.n_pm + 1 + 1) / 2 * 2);
  float* s_b1 = reinterpret_cast<float*>(s_frag + HW * TG::NFRAG * 32);
  float* s_xch = s_b1 + HW * TG::HWARP;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kPerWarp = GRAD ? TG::WARP_F : 0;
  float* s_warp = s_xch + TG::XCH_F + (size_t)wl * kPerWarp;
  float* s_stage = s_warp;
  float* s_zs = s_warp + 2 * TG::REC;
  float* s_ks = s_zs + TG::HTILE;
  uint64_t* bar = s_bar[wl];
  const int g = lane >> 2, t = lane & 3;
  const long long total_groups = gridDim.x;
  const int S = P.substeps;
  const double invS = 1.0 / (double)S;
  constexpr uint32_t kRecBytes = TG::REC * 4;
  const long long warp_global = (long long)blockIdx.x * HW + wl;
  float* __restrict__ stash = GRAD ? reinterpret_cast<float*>(P.stash) + warp_global * P.stash_rows_per_warp * 32 : nullptr;
  int xpar = 0;
  const bool lead = wl == 0;                       // the lead warp owns the loss, the uhat gradient and the b2 gradient

  for (long long k = threadIdx.x; k <= P.n_pm; k += blockDim.x) s_red[k] = 0.0;
  const double* __restrict__ pm = P.theta + (long long)D * P.model_col0[1];        // single model: [uhat | pm]
  auto W1 = [&](int hid, int slot) -> float { return (hid < P.H) ? (float)pm[P.off_w1 + hid + (long long)P.H * slot] : 0.f; };
  auto W2 = [&](int o, int hid) -> float { return (hid < P.H) ? (float)pm[P.off_w2 + o + (long long)P.nn_out * hid] : 0.f; };
  // pre-split B fragments: [warp][product][tile][lane] -> (b0hi, b1hi, b0lo, b1lo)
  for (int idx = threadIdx.x; idx < HW * TG::NFRAG * 32; idx += blockDim.x) {
    const int l = idx & 31, f = (idx >> 5) % TG::NFRAG, w = idx / (32 * TG::NFRAG);
    const int gg = l >> 2, tt = l & 3, prod = f / HTW, j = f % HTW, hb = w * TG::HWARP + 8 * j;
    float b0, b1;
    if (prod == 0) { b0 = W1(hb + gg, tt); b1 = W1(hb + gg, tt + 4); }                                            // z = x W1^T
    else if (prod == 1) { b0 = W2(tf_omap(gg), hb + tf_kmap(tt)); b1 = W2(tf_omap(gg), hb + tf_kmap(tt + 4)); }   // y = h W2^T
    else if (prod == 2) { b0 = W2(tt, hb + gg); b1 = W2(tt + 4, hb + gg); }                                       // h-bar = k-bar W2
    else { b0 = W1(hb + tf_kmap(tt), tf_omap(gg)); b1 = W1(hb + tf_kmap(tt + 4), tf_omap(gg)); }                  // x-bar = z-bar W1
    uint4 q;
    split_tf32(b0, q.x, q.z); split_tf32(b1, q.y, q.w);
    s_frag[idx] = q;
  }
  for (int idx = threadIdx.x; idx < HW * TG::HWARP; idx += blockDim.x) s_b1[idx] = (idx < P.H) ? (float)pm[P.off_b1 + idx] : 0.f;
  if (GRAD && lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  uint32_t bar_phase = 0;
  __syncthreads();
  const uint4* __restrict__ fr = s_frag + (size_t)wl * TG::NFRAG * 32 + lane;      // + f * 32

  float b2own[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) b2own[j] = (float)pm[P.off_b2 + t + 4 * j];

  // gradient accumulators: C fragments with M = hidden (two m-tiles of 16 per warp when HTW = 4), N = slots / outputs
  constexpr int MT = HTW / 2;
  static_assert(HTW % 2 == 0, "two n-tiles form one 16-row m-tile of the gradient products");
  float GW1[MT][4], GW2[MT][4], gb1[HTW][2], gb2[2] = {0.f, 0.f};
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int q = 0; q < 4; ++q) { GW1[m][q] = 0.f; GW2[m][q] = 0.f; }
#pragma unroll
  for (int j = 0; j < HTW; ++j) { gb1[j][0] = 0.f; gb1[j][1] = 0.f; }
  double loss = 0.0;

  const double* __restrict__ uh = P.uhat_x ? P.uhat_x : P.theta;
  double* __restrict__ guh = P.guhat_x ? P.guhat_x : P.grad;
  const bool multi = P.loss_kind == UDE_LOSS_MULTI_SHOOT;

  // one RHS evaluation: xs[half][slot e] (this lane's slots t + 4e of segments g and g + 8) -> h (C layout), y on my slots
  auto stage_forward = [&](const float (&xs)[2][2], float (&h)[HTW][4], float (&y)[2][2]) {
    const float a[4] = {xs[0][0], xs[1][0], xs[0][1], xs[1][1]};
    uint32_t ah[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
    for (int j = 0; j < HTW; ++j) {
      const float2 bb = *reinterpret_cast<const float2*>(s_b1 + wl * TG::HWARP + 8 * j + 2 * t);
      h[j][0] = bb.x; h[j][1] = bb.y; h[j][2] = bb.x; h[j][3] = bb.y;
      const uint4 b = fr[(0 * HTW + j) * 32];
      mma_tf32(h[j], al, b.x, b.y);
      mma_tf32(h[j], ah, b.z, b.w);
      mma_tf32(h[j], ah, b.x, b.y);
    }
#pragma unroll
    for (int j = 0; j < HTW; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) h[j][q] = tf_tanhf(h[j][q]);
    float yc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < HTW; ++j) {
      const float ak[4] = {h[j][0], h[j][2], h[j][1], h[j][3]};
      mma3(yc, ak, fr[(1 * HTW + j) * 32]);
    }
    tf_xchg(yc, s_xch, xpar, wl, lane);
    y[0][0] = yc[0] + b2own[0]; y[0][1] = yc[1] + b2own[1];      // C columns 2t, 2t+1 -> outputs t, t+4; rows g, g+8
    y[1][0] = yc[2] + b2own[0]; y[1][1] = yc[3] + b2own[1];
  };

  Seg nxt[2];
#pragma unroll
  for (int tl = 0; tl < 2; ++tl) {
    const long long t0 = P.task_begin + blockIdx.x;
    nxt[tl] = P.segs[(t0 < P.n_tasks ? t0 : P.n_tasks - 1) * 16 + 8 * tl + g];
  }

  for (long long task = P.task_begin + blockIdx.x; task < P.n_tasks; task += total_groups) {
    Seg sg[2];
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) sg[tl] = nxt[tl];
    {
      const long long tn = task + total_groups;
#pragma unroll
      for (int tl = 0; tl < 2; ++tl) nxt[tl] = P.segs[(tn < P.n_tasks ? tn : task) * 16 + 8 * tl + g];
    }
    bool is_null[2], has_prev[2];
    int c0[2], L[2];
    double wpt[2];
    int maxL = 0, pre = 0, first_scored = 0;
    bool from_theta = true;
    double eps = 0.0;
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
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
      // every warp of the block works on the same 16 segments: the block-wide maximum is the warp-wide one
      maxL = (int)__reduce_max_sync(0xffffffffu, (unsigned)maxL);
      from_theta = multi || P.shoot_from_theta;
      first_scored = multi ? 0 : P.shoot_first_scored;
      eps = multi ? P.preroll_eps : 0.0;
      pre = eps > 0.0 ? 1 : 0;
    }
    const int nsteps = pre + maxL * S;

    float u[2][KSU];
    double tq0[2], tq1[2], dq[2][KSU];
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
      tq0[tl] = P.times[c0[tl]];
      tq1[tl] = P.times[c0[tl] + min(1, L[tl])];
#pragma unroll
      for (int e = 0; e < KSU; ++e) {
        const int s = t + 4 * e;
        u[tl][e] = (float)(from_theta ? uh[(long long)D * c0[tl] + s] : P.data[(long long)D * c0[tl] + s]);
        dq[tl][e] = (LK == LK_TRAJ) ? P.data[(long long)D * c0[tl] + s] : 0.0;
      }
    }

    // ------------------------------------------ forward sweep ------------------------------------------------------
    {
      int i = pre ? -1 : 0, n = pre ? S - 1 : 0;
      float hs[2];
#pragma unroll
      for (int tl = 0; tl < 2; ++tl) hs[tl] = is_null[tl] ? 0.f : (float)eps;
      for (int k = 0; k < nsteps; ++k) {
        if (k >= pre && n == 0) {
#pragma unroll
          for (int tl = 0; tl < 2; ++tl) {
            if (LK == LK_TRAJ) {
              const double w = (i <= L[tl] && i >= first_scored) ? wpt[tl] : 0.0;
#pragma unroll
              for (int e = 0; e < KSU; ++e) {
                const double res = dq[tl][e] - (double)u[tl][e];
                if (lead) loss = fma(w * res, res, loss);
                if (!GRAD && lead && P.traj_out != nullptr && i <= L[tl] && !is_null[tl]) P.traj_out[(long long)D * (c0[tl] + i) + t + 4 * e] = (double)u[tl][e];
              }
            }
            hs[tl] = (i < L[tl]) ? (float)((tq1[tl] - tq0[tl]) * invS) : 0.f;
            tq0[tl] = tq1[tl];
            tq1[tl] = ld_pf(P.times + c0[tl] + min(i + 2, L[tl]));
            if (LK == LK_TRAJ) {
#pragma unroll
              for (int e = 0; e < KSU; ++e) dq[tl][e] = ld_pf(P.data + (long long)D * (c0[tl] + min(i + 1, L[tl])) + t + 4 * e);
            }
          }
        }
        float kst[2][T::S][KSU], acc[2][KSU];
#pragma unroll
        for (int tl = 0; tl < 2; ++tl)
#pragma unroll
          for (int e = 0; e < KSU; ++e) acc[tl][e] = 0.f;
#pragma unroll
        for (int st = 0; st < T::S; ++st) {
          float xs[2][2], h[HTW][4], y[2][2];
#pragma unroll
          for (int tl = 0; tl < 2; ++tl)
#pragma unroll
            for (int e = 0; e < KSU; ++e) {
              float a = 0.f;
#pragma unroll
              for (int l = 0; l < st; ++l) if (T::a(st, l) != 0.0) a = fmaf((float)T::a(st, l), kst[tl][l][e], a);
              xs[tl][e] = (st == 0) ? u[tl][e] : fmaf(hs[tl], a, u[tl][e]);
            }
          stage_forward(xs, h, y);
          if (GRAD) {
            const int rec = k * T::S + st;
            float* sbuf = s_stage + (rec & 1) * TG::REC;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < HTW; ++j) {
              *reinterpret_cast<float2*>(sbuf + g * TG::HS + 8 * j + 2 * t) = make_float2(h[j][0], h[j][1]);
              *reinterpret_cast<float2*>(sbuf + (g + 8) * TG::HS + 8 * j + 2 * t) = make_float2(h[j][2], h[j][3]);
            }
#pragma unroll
            for (int tl = 0; tl < 2; ++tl)
#pragma unroll
              for (int e = 0; e < KSU; ++e) sbuf[TG::HTILE + (g + 8 * tl) * TG::XS + t + 4 * e] = xs[tl][e];
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_store(stash + (long long)rec * TG::REC, sbuf, kRecBytes); bulk_commit(); }
          }
#pragma unroll
          for (int tl = 0; tl < 2; ++tl)
#pragma unroll
            for (int e = 0; e < KSU; ++e) { kst[tl][st][e] = y[tl][e]; acc[tl][e] = fmaf((float)T::b(st), y[tl][e], acc[tl][e]); }
        }
#pragma unroll
        for (int tl = 0; tl < 2; ++tl)
#pragma unroll
          for (int e = 0; e < KSU; ++e) u[tl][e] = fmaf(hs[tl], acc[tl][e], u[tl][e]);
        if (++n == S) { n = 0; ++i; }
      }
      if (GRAD) {
        if (lane == 0) {
          bulk_wait_all<0>();
          if (nsteps > 0) {
            const int rec = nsteps * T::S - 1, b = rec & 1;
            mbar_expect_tx(&bar[b], kRecBytes);
            bulk_load(s_stage + b * TG::REC, stash + (long long)rec * TG::REC, kRecBytes, &bar[b]);
          }
        }
        __syncwarp();
      }
    }
    // the next task's first columns: bring their lines into L2 now (its segment entries were requested at this task's start)
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
      const long long cn = (LK == LK_STATE && !(nxt[tl].flags & SEG_NO_PREV)) ? nxt[tl].col - 1 : nxt[tl].col;
      if (t == 0) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.times + cn));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.data + (long long)D * cn));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(uh + (long long)D * cn));
        if (LK == LK_STATE) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(P.data + (long long)D * (cn + 1)));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(uh + (long long)D * (cn + 1)));
        }
      }
    }
    // ------------------------------------------ loss at the last point ------------------------------------------------
    float lam[2][KSU];
    double gc[2][KSU], gp[2][KSU], rt_lo[2], rt_hi[2], dr[2][KSU];
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
#pragma unroll
      for (int e = 0; e < KSU; ++e) { lam[tl][e] = 0.f; gc[tl][e] = 0.0; gp[tl][e] = 0.0; dr[tl][e] = 0.0; }
      const int ir = max(maxL - 1, 0);
      rt_lo[tl] = P.times[c0[tl] + min(ir, L[tl])];
      rt_hi[tl] = P.times[c0[tl] + min(ir + 1, L[tl])];
      if (LK == LK_STATE) {
        const int col = sg[tl].col;
        const double dt = P.times[col] - P.times[c0[tl]];
        const double wsc = (L[tl] == 1) ? P.proc_scale[0] * (P.proc_inv_dt ? 1.0 / dt : 1.0) : 0.0;
        const double os = is_null[tl] ? 0.0 : P.obs_scale[0];
        const bool obs_first = (sg[tl].flags & SEG_OBS_FIRST) && has_prev[tl];
        double rown[2], eown[2], fown[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          rown[j] = uh[(long long)D * col + t + 4 * j] - (double)u[tl][j];
          eown[j] = P.data[(long long)D * col + t + 4 * j] - uh[(long long)D * col + t + 4 * j];
          fown[j] = obs_first ? P.data[(long long)D * c0[tl] + t + 4 * j] - uh[(long long)D * c0[tl] + t + 4 * j] : 0.0;
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
        if (t == 0 && lead) loss += wsc * q1 + os * (q2 + q3);
#pragma unroll
        for (int e = 0; e < KSU; ++e) {
#pragma unroll
          for (int b = 0; b < D; ++b) if (b == t + 4 * e) {
            lam[tl][e] = (float)(-wsc * sym[b]); gc[tl][e] = wsc * sym[b] - os * es[b]; gp[tl][e] = -os * fs[b];
          }
        }
        if (!GRAD && lead && P.traj_out != nullptr && !is_null[tl]) {
#pragma unroll
          for (int e = 0; e < KSU; ++e) {
            if (L[tl] == 1) P.traj_out[(long long)D * col + t + 4 * e] = (double)u[tl][e];
            if (!has_prev[tl] || (sg[tl].flags & SEG_OBS_FIRST)) P.traj_out[(long long)D * c0[tl] + t + 4 * e] = uh[(long long)D * c0[tl] + t + 4 * e];
          }
        }
      } else {
        const double w = (maxL <= L[tl] && maxL >= first_scored) ? wpt[tl] : 0.0;
#pragma unroll
        for (int e = 0; e < KSU; ++e) {
          const double res = dq[tl][e] - (double)u[tl][e];
          if (lead) loss = fma(w * res, res, loss);
          lam[tl][e] = (float)(-2.0 * w * res);
          if (!GRAD && lead && P.traj_out != nullptr && maxL <= L[tl] && !is_null[tl] && !(multi && L[tl] == P.pred_length))
            P.traj_out[(long long)D * (c0[tl] + L[tl]) + t + 4 * e] = (double)u[tl][e];
        }
      }
    }
    // ------------------------------------------ reverse sweep ----------------------------------------------------------
    if (GRAD) {
      int i = maxL - 1, n = S - 1;
      float hs[2] = {0.f, 0.f};
      for (int k = nsteps - 1; k >= 0; --k) {
        if (k >= pre) {
          if (n == S - 1) {
#pragma unroll
            for (int tl = 0; tl < 2; ++tl) {
              hs[tl] = (i < L[tl]) ? (float)((rt_hi[tl] - rt_lo[tl]) * invS) : 0.f;
              if (LK == LK_TRAJ) {
#pragma unroll
                for (int e = 0; e < KSU; ++e) dr[tl][e] = ld_pf(P.data + (long long)D * (c0[tl] + min(i, L[tl])) + t + 4 * e);
              }
              rt_hi[tl] = rt_lo[tl];
              rt_lo[tl] = ld_pf(P.times + c0[tl] + min(max(i - 1, 0), L[tl]));
            }
          }
        } else {
#pragma unroll
          for (int tl = 0; tl < 2; ++tl) hs[tl] = is_null[tl] ? 0.f : (float)eps;
        }
        float Ub[2][T::S][KSU], lamacc[2][KSU], U0[2][KSU];
#pragma unroll
        for (int tl = 0; tl < 2; ++tl)
#pragma unroll
          for (int e = 0; e < KSU; ++e) { lamacc[tl][e] = 0.f; U0[tl][e] = 0.f; }
#pragma unroll
        for (int ii = 0; ii < T::S; ++ii) {
          const int st = T::S - 1 - ii;
          const int rec = k * T::S + st, b = rec & 1;
          if (rec > 0 && lane == 0) {
            mbar_expect_tx(&bar[b ^ 1], kRecBytes);
            bulk_load(s_stage + (b ^ 1) * TG::REC, stash + (long long)(rec - 1) * TG::REC, kRecBytes, &bar[b ^ 1]);
          }
          mbar_wait(&bar[b], (bar_phase >> b) & 1u);
          bar_phase ^= (1u << b);
          const float* __restrict__ sbuf = s_stage + b * TG::REC;
          float h[HTW][4], kb[2][2];
#pragma unroll
          for (int j = 0; j < HTW; ++j) {
            const float2 lo = *reinterpret_cast<const float2*>(sbuf + g * TG::HS + 8 * j + 2 * t);
            const float2 hi = *reinterpret_cast<const float2*>(sbuf + (g + 8) * TG::HS + 8 * j + 2 * t);
            h[j][0] = lo.x; h[j][1] = lo.y; h[j][2] = hi.x; h[j][3] = hi.y;
          }
#pragma unroll
          for (int tl = 0; tl < 2; ++tl) {
            if (st == 0) {
#pragma unroll
              for (int e = 0; e < KSU; ++e) U0[tl][e] = sbuf[TG::HTILE + (g + 8 * tl) * TG::XS + t + 4 * e];
            }
#pragma unroll
            for (int e = 0; e < KSU; ++e) {
              float a = (float)T::b(st) * lam[tl][e];
#pragma unroll
              for (int l = st + 1; l < T::S; ++l) if (T::a(l, st) != 0.0) a = fmaf((float)T::a(l, st), Ub[tl][l][e], a);
              kb[tl][e] = hs[tl] * a;
            }
          }
          // h-bar = k-bar W2 (one k-step over the eight outputs), then z-bar
          float zb[HTW][4];
          {
            const float ak[4] = {kb[0][0], kb[1][0], kb[0][1], kb[1][1]};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) split_tf32(ak[q], ah[q], al[q]);
#pragma unroll
            for (int j = 0; j < HTW; ++j) {
              zb[j][0] = zb[j][1] = zb[j][2] = zb[j][3] = 0.f;
              const uint4 bq = fr[(2 * HTW + j) * 32];
              mma_tf32(zb[j], al, bq.x, bq.y);
              mma_tf32(zb[j], ah, bq.z, bq.w);
              mma_tf32(zb[j], ah, bq.x, bq.y);
            }
          }
#pragma unroll
          for (int j = 0; j < HTW; ++j) {
#pragma unroll
            for (int q = 0; q < 4; ++q) zb[j][q] *= fmaf(-h[j][q], h[j][q], 1.f);
            *reinterpret_cast<float2*>(s_zs + g * TG::HS + 8 * j + 2 * t) = make_float2(zb[j][0], zb[j][1]);
            *reinterpret_cast<float2*>(s_zs + (g + 8) * TG::HS + 8 * j + 2 * t) = make_float2(zb[j][2], zb[j][3]);
            gb1[j][0] += zb[j][0] + zb[j][2]; gb1[j][1] += zb[j][1] + zb[j][3];
          }
#pragma unroll
          for (int tl = 0; tl < 2; ++tl)
#pragma unroll
            for (int e = 0; e < KSU; ++e) s_ks[(g + 8 * tl) * TG::XS + t + 4 * e] = kb[tl][e];
          if (lead) { gb2[0] += kb[0][0] + kb[1][0]; gb2[1] += kb[0][1] + kb[1][1]; }
          // x-bar = z-bar W1 on my slots (partial over my hidden units, then summed over the warps)
          float xc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < HTW; ++j) {
            const float az[4] = {zb[j][0], zb[j][2], zb[j][1], zb[j][3]};
            mma3(xc, az, fr[(3 * HTW + j) * 32]);
          }
          tf_xchg(xc, s_xch, xpar, wl, lane);
          Ub[0][st][0] = xc[0]; Ub[0][st][1] = xc[1]; Ub[1][st][0] = xc[2]; Ub[1][st][1] = xc[3];
#pragma unroll
          for (int tl = 0; tl < 2; ++tl)
#pragma unroll
            for (int e = 0; e < KSU; ++e) lamacc[tl][e] += Ub[tl][st][e];
          __syncwarp();                       // z-bar / k-bar tiles complete
          // weight gradients: M = my hidden units (m-tiles of 16), N = slots / outputs, K = the 16 segments (two k-steps).
          // The tensor core's FP32 accumulator does not round to nearest: a C fragment that lives across thousands of MMAs
          // picks up a BIAS of ~n 2^-24 (measured: 4e-5 on the C5 weight gradients at 2e5 segments, growing with the problem).
          // So each stage's products go into fresh accumulators and are added to the running sums with ordinary FADDs.
          float p1[MT][4], p2[MT][4];
#pragma unroll
          for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int q = 0; q < 4; ++q) { p1[m][q] = 0.f; p2[m][q] = 0.f; }
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const int s0 = 8 * ks + t, s1 = s0 + 4;
            const float bx0 = sbuf[TG::HTILE + s0 * TG::XS + g], bx1 = sbuf[TG::HTILE + s1 * TG::XS + g];     // x[seg][slot g]
            const float bk0 = s_ks[s0 * TG::XS + g], bk1 = s_ks[s1 * TG::XS + g];                             // k-bar[seg][out g]
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              const int h0 = 16 * m + g, h1 = h0 + 8;
              const float az[4] = {s_zs[s0 * TG::HS + h0], s_zs[s0 * TG::HS + h1], s_zs[s1 * TG::HS + h0], s_zs[s1 * TG::HS + h1]};
              const float ahh[4] = {sbuf[s0 * TG::HS + h0], sbuf[s0 * TG::HS + h1], sbuf[s1 * TG::HS + h0], sbuf[s1 * TG::HS + h1]};
              mma3_ff(p1[m], az, bx0, bx1);
              mma3_ff(p2[m], ahh, bk0, bk1);
            }
          }
#pragma unroll
          for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int q = 0; q < 4; ++q) { GW1[m][q] += p1[m][q]; GW2[m][q] += p2[m][q]; }
          __syncwarp();                       // everyone is done with buffer b and the scratch tiles
          if (P.stash_discard) {
            for (int ln = lane; ln < (int)(kRecBytes / 128); ln += 32)
              asm volatile("discard.global.L2 [%0], 128;" ::"l"(stash + (long long)rec * TG::REC + ln * 32) : "memory");
          }
        }
#pragma unroll
        for (int tl = 0; tl < 2; ++tl)
#pragma unroll
          for (int e = 0; e < KSU; ++e) lam[tl][e] += lamacc[tl][e];
        if (k >= pre) {
          if (n == 0) {
            if (LK == LK_TRAJ) {
#pragma unroll
              for (int tl = 0; tl < 2; ++tl) {
                const double w = (i <= L[tl] && i >= first_scored) ? wpt[tl] : 0.0;
#pragma unroll
                for (int e = 0; e < KSU; ++e) lam[tl][e] = (float)fma(-2.0 * w, dr[tl][e] - (double)U0[tl][e], (double)lam[tl][e]);
              }
            }
            --i; n = S - 1;
          } else --n;
        }
      }
      if (lead) {
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
          if (!is_null[tl]) {
#pragma unroll
            for (int e = 0; e < KSU; ++e) {
              if (LK == LK_STATE) {
                atomicAdd(guh + (long long)D * sg[tl].col + t + 4 * e, gc[tl][e]);
                if (has_prev[tl]) atomicAdd(guh + (long long)D * c0[tl] + t + 4 * e, gp[tl][e] + (double)lam[tl][e]);
              } else if (from_theta) {
                if (P.guhat_x) guh[(long long)D * c0[tl] + t + 4 * e] = (double)lam[tl][e];
                else atomicAdd(guh + (long long)D * c0[tl] + t + 4 * e, (double)lam[tl][e]);
              }
            }
          }
        }
      }
    }
  }

  // ---- flush ------------------------------------------------------------------------------------------------------
  {
    double l = loss;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
    // bias gradients: sum over the 8 row groups g (lane bits 2..4)
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
      gb2[0] += __shfl_xor_sync(0xffffffffu, gb2[0], off); gb2[1] += __shfl_xor_sync(0xffffffffu, gb2[1], off);
#pragma unroll
      for (int j = 0; j < HTW; ++j) { gb1[j][0] += __shfl_xor_sync(0xffffffffu, gb1[j][0], off); gb1[j][1] += __shfl_xor_sync(0xffffffffu, gb1[j][1], off); }
    }
    // every warp owns disjoint hidden units: no ordering needed between the warps, except for gb2 / loss (lead only)
    if (GRAD) {
#pragma unroll
      for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int hid = wl * TG::HWARP + 16 * m + g + 8 * (q >> 1);     // C rows g, g+8
          const int col = 2 * t + (q & 1);                                // slot / output
          if (hid < P.H) {
            s_red[P.off_w1 + hid + (long long)P.H * col] += (double)GW1[m][q];
            s_red[P.off_w2 + col + (long long)P.nn_out * hid] += (double)GW2[m][q];
          }
        }
      }
      if (g == 0) {
#pragma unroll
        for (int j = 0; j < HTW; ++j)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int hid = wl * TG::HWARP + 8 * j + 2 * t + q;
            if (hid < P.H) s_red[P.off_b1 + hid] += (double)gb1[j][q];
          }
        if (lead) { s_red[P.off_b2 + t] += (double)gb2[0]; s_red[P.off_b2 + t + 4] += (double)gb2[1]; }
      }
    }
    if (lead && lane == 0) s_red[P.n_pm] += l;
    __syncthreads();
    double* row = P.block_part + (long long)blockIdx.x * (P.n_pm + 1);
    for (long long k = threadIdx.x; k <= P.n_pm; k += blockDim.x) row[k] = s_red[k];
  }
}

}  // namespace ude
