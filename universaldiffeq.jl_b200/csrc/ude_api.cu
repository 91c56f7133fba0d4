// libude_cuda.so -- C ABI (include/ude_cuda.h) over the sm_100a kernels.  There is no CPU fallback: every entry
// point that computes fails with UDE_ERR_NO_DEVICE / UDE_ERR_CUDA when no Blackwell GPU is usable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ude_cuda.h"
#include "ude_device.cuh"
#include "ude_registry.h"

namespace ude {
std::vector<KernelEntry>& kernel_registry() {
  static std::vector<KernelEntry> r;
  return r;
}
}  // namespace ude

using ude::DevProblem;
using ude::KernelEntry;
using ude::Seg;

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  cudaError_t reserve(size_t count) {
    if (count <= n && p) return cudaSuccess;
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t upload(const T* src, size_t count, cudaStream_t s) {
    cudaError_t e = reserve(count);
    if (e != cudaSuccess) return e;
    if (count == 0) return cudaSuccess;
    return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

// ---- NCCL through dlopen: the single-GPU path has no NCCL dependency --------------------------------------------
struct Id128 { char b[128]; };   // ncclUniqueId, passed BY VALUE to ncclCommInitRank
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

std::mutex g_nccl_mutex;

bool nccl_load(std::string& err) {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);      // handles on different host threads may get here together
  if (g_nccl.lib && g_nccl.AllReduce) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) { g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) { err = "NCCL symbols missing"; return false; }
  return true;
}

}  // namespace

namespace { struct ude_ukf_plan; }

struct ude_handle {
  ude_desc desc{};
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  mutable std::string err;
  // host copies of the index arrays
  int64_t n_series = 0, n_cols = 0, n_models = 1, n_theta = 0;
  std::vector<int64_t> starts, lengths, model_of_series, theta_base, model_col0, model_series0;
  std::vector<uint8_t> skip;
  std::vector<double> obs_scale, proc_scale, shoot_weight, reg_weight;
  bool have_data = false, have_cov = false, have_targets = false, dirty = true;
  bool sparse_uhat_io = false;          // UDE_OPT_SPARSE_UHAT_IO
  double* h_stage = nullptr; size_t h_stage_n = 0;       // pinned [loss | process-model gradient] staging: one D2H copy per call
  bool ukf_causal = false;              // UDE_OPT_UKF_CAUSAL_INTERVAL
  int64_t last_h2d = 0, last_d2h = 0, n_segments = 0;
  // device buffers
  DevBuf<long long> d_starts, d_lengths, d_theta_base, d_model_col0, d_model_series0, d_knot_off;
  DevBuf<double> d_times, d_data, d_obs_scale, d_proc_scale, d_shoot_weight, d_reg_weight;
  DevBuf<double> d_knot_t, d_knot_x, d_knot_inv, d_dm_u, d_dm_dudt;
  DevBuf<Seg> d_segs;
  DevBuf<double> d_theta, d_grad, d_loss, d_block_part, d_model_loss, d_stash, d_traj, d_red, d_adam_m, d_adam_v;
  // launch plan
  const KernelEntry* k_grad = nullptr;
  const KernelEntry* k_fwd = nullptr;
  int64_t n_tasks = 0, steps_per_eval = 0, max_steps_per_task = 0;
  int grid_grad = 0, grid_fwd = 0;
  long long stash_rows_per_warp = 0;
  size_t smem_bytes = 0, smem_bytes_fwd = 0;
  int adam_used_graph = 0;
  std::vector<int64_t> model_task0;     // many-model handles: tasks of model m are [model_task0[m], model_task0[m + 1])
  ude_ukf_plan* ukf = nullptr;          // unscented-Kalman-filter loss (ude_ukf.cuh), built on first use
  // multi-GPU
  void* comm = nullptr;
  int n_ranks = 1, rank = 0;
  // fused finalize + allreduce over peer memory (ude_peer_export / ude_peer_attach)
  void* x_local = nullptr;              // this rank's exchange buffer (cudaMalloc; exported by IPC handle)
  size_t x_bytes = 0;
  int x_ranks = 0, x_blocks = 0;
  long long x_stride = 0;
  void* x_peer[UDE_MAX_PEERS] = {};     // every rank's buffer mapped here (x_peer[rank] == x_local)
  bool x_opened[UDE_MAX_PEERS] = {};    // mapped with cudaIpcOpenMemHandle (to be closed)
  bool x_attached = false;
  // host-buffer path: series chunks whose H2D / kernel / D2H overlap (built with the plan)
  // single model: up_* / dn_* are uhat COLUMN ranges.  Many models: a chunk is a run of whole models [m_lo, m_hi) and
  // up_* / dn_* are the (identical) ranges of theta / gradient ENTRIES of those models.
  struct Chunk { int64_t task_begin, task_end, up_lo, up_hi, dn_lo, dn_hi, m_lo = 0, m_hi = 0; };
  std::vector<Chunk> chunks;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev_up, ev_done;
  cudaEvent_t ev_start = nullptr;
  // optional kernel timing (ude_profile_begin / ude_profile_read)
  std::vector<cudaEvent_t> ev0, ev1;
  int ev_used = 0, ev_cap = 0;
};

namespace {

int fail(const ude_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}
#define UDE_CUDA_TRY(h, expr)                                                                         \
  do { cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail((h), UDE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

// ---- reductions that follow the segment kernel --------------------------------------------------------------------
// single model: out[k] = sum_b block_part[b][k] (+ L2 regulariser); loss likewise.  Fixed summation order.
__global__ void ude_finalize_single(const double* __restrict__ block_part, int n_blocks, long long n_pm,
                                    const double* __restrict__ pm, long long off_w1, long long n_w1, long long off_w2,
                                    long long n_w2, long long series_lo, long long series_hi, const double* __restrict__ reg_weight,
                                    double* __restrict__ grad_pm, double* __restrict__ loss_out, int accumulate_series) {
  __shared__ double sh[8][33];
  const int x = threadIdx.x, y = threadIdx.y;
  const long long k = (long long)blockIdx.x * 32 + x;
  const double rw = reg_weight ? reg_weight[0] : 0.0;
  double acc = 0.0;
  if (k <= n_pm) {
    for (int b = y; b < n_blocks; b += 8) acc += block_part[(long long)b * (n_pm + 1) + k];
    if (k == n_pm && rw != 0.0) {   // the loss slot: add rw * (|W1|^2 + |W2|^2), partitioned over y
      double r = 0.0;
      for (long long i = y; i < n_w1; i += 8) { double v = pm[off_w1 + i]; r = fma(v, v, r); }
      for (long long i = y; i < n_w2; i += 8) { double v = pm[off_w2 + i]; r = fma(v, v, r); }
      acc = fma(rw, r, acc);
    }
  }
  sh[y][x] = acc;
  __syncthreads();
  if (y == 0 && k <= n_pm) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i][x];
    if (k == n_pm) { loss_out[0] = s; }
    else {
      const bool isw = (k >= off_w1 && k < off_w1 + n_w1) || (k >= off_w2 && k < off_w2 + n_w2);
      if (isw && rw != 0.0) s = fma(2.0 * rw, pm[k], s);
      if (k >= series_lo && k < series_hi) { if (accumulate_series) grad_pm[k] += s; }   // written by atomics in the segment kernel
      else grad_pm[k] = s;
    }
  }
}

// many models: loss[m] = model_loss[m] + reg; grad_pm += 2 rw W.  One block per model.
__global__ void ude_finalize_multi(const double* __restrict__ theta, double* __restrict__ grad, const double* __restrict__ model_loss,
                                   const long long* __restrict__ theta_base, const long long* __restrict__ model_col0, int d,
                                   long long off_w1, long long n_w1, long long off_w2, long long n_w2,
                                   const double* __restrict__ reg_weight, double* __restrict__ loss_out) {
  const int m = blockIdx.x;
  const long long base = theta_base[m] + (long long)d * (model_col0[m + 1] - model_col0[m]);
  const double rw = reg_weight ? reg_weight[m] : 0.0;
  double r = 0.0;
  if (rw != 0.0) {
    for (long long i = threadIdx.x; i < n_w1 + n_w2; i += blockDim.x) {
      const long long k = i < n_w1 ? off_w1 + i : off_w2 + (i - n_w1);
      const double v = theta[base + k];
      r = fma(v, v, r);
      if (grad) grad[base + k] = fma(2.0 * rw, v, grad[base + k]);
    }
  }
  __shared__ double sh[128];
  sh[threadIdx.x] = r;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) loss_out[m] = fma(rw, sh[0], model_loss[m]);
}

// red = [loss | process-model gradient]; the per-series tail was produced by atomics and is left untouched
__global__ void ude_scatter_reduced(const double* __restrict__ red, long long n_pm, long long series_lo, long long series_hi,
                                    double* __restrict__ grad_pm, double* __restrict__ loss_out) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_pm) { if (grad_pm && !(k >= series_lo && k < series_hi)) grad_pm[k] = red[1 + k]; }
  else if (k == n_pm) loss_out[0] = red[0];
}


// ---- finalize fused with the cross-GPU exchange ("fuse with the grid-level reduction epilogue") --------
// One launch does what ude_finalize_single + ncclAllReduce + ude_scatter_reduced do in three: each block reduces its 32
// columns of the per-block rows (same fixed order), PUSHES the partial sums straight into every rank's exchange buffer
// with peer stores over NVLink, raises a per-(sender, block) flag, waits for the same block of every other rank and
// adds the n_ranks contributions in rank order -- so all ranks hold bit-identical sums.
// Exchange buffer of a rank (x): slots  [2 parities][n_ranks senders][stride]  doubles,
//                                flags  [2 parities][n_ranks senders][n_blocks] epoch numbers,
//                                epoch  [n_blocks]  (this rank's own call counter, device-resident so that the launch
//                                                    has no per-call argument and can sit in a captured graph),
//                                status [1]         (non-zero: a wait ran out of time).
// Parity double-buffering is enough: a sender reaches epoch e+2 only after it saw this rank's flag of epoch e+1, which
// this rank raises after its epoch-e launch has finished reading.
struct PeerTable {
  double* slot[UDE_MAX_PEERS];
  unsigned long long* flag[UDE_MAX_PEERS];
  unsigned long long* epoch;       // local
  unsigned long long* status;      // local
  long long stride;
  int n_ranks, rank, n_blocks;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void ude_finalize_exchange(const double* __restrict__ block_part, int n_blocks, long long n_pm, long long n_shared,
                                      const double* __restrict__ pm, long long off_w1, long long n_w1, long long off_w2,
                                      long long n_w2, const double* __restrict__ reg_weight,
                                      double* __restrict__ grad_pm, double* __restrict__ loss_out, PeerTable T,
                                      unsigned long long timeout_ns) {
  // exchange index j: j < n_shared -> process-model entry j; j == n_shared -> the loss (column n_pm of the block rows).
  // The rank-local per-series tail [n_shared, n_pm) was produced by atomics and takes no part (ranks may own different
  // numbers of series, so the grid covers the exchanged entries only and is the same on every rank).
  __shared__ double sh[8][33];
  const int x = threadIdx.x, y = threadIdx.y;
  const long long j = (long long)blockIdx.x * 32 + x;
  const bool live = j <= n_shared;
  const long long k = j < n_shared ? j : n_pm;
  const double rw = reg_weight ? reg_weight[0] : 0.0;
  const unsigned long long epoch = T.epoch[blockIdx.x] + 1;
  const int par = (int)(epoch & 1);
  double acc = 0.0;
  if (live) {
    for (int b = y; b < n_blocks; b += 8) acc += block_part[(long long)b * (n_pm + 1) + k];
    if (k == n_pm && rw != 0.0) {
      double r = 0.0;
      for (long long i = y; i < n_w1; i += 8) { double v = pm[off_w1 + i]; r = fma(v, v, r); }
      for (long long i = y; i < n_w2; i += 8) { double v = pm[off_w2 + i]; r = fma(v, v, r); }
      acc = fma(rw, r, acc);
    }
  }
  sh[y][x] = acc;
  __syncthreads();
  if (y == 0 && live) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i][x];
    if (k < n_pm) {
      const bool isw = (k >= off_w1 && k < off_w1 + n_w1) || (k >= off_w2 && k < off_w2 + n_w2);
      if (isw && rw != 0.0) s = fma(2.0 * rw, pm[k], s);
    }
    const long long at = ((long long)par * T.n_ranks + T.rank) * T.stride + j;
    for (int r = 0; r < T.n_ranks; ++r) T.slot[(r + T.rank) % T.n_ranks][at] = s;    // peer stores, staggered start
    __threadfence_system();
  }
  __syncthreads();
  const long long fat = ((long long)par * T.n_ranks) * T.n_blocks + blockIdx.x;
  if (y == 0 && x < T.n_ranks) {
    // lane x raises this rank's flag at rank x, then waits for rank x's flag here
    st_release_sys(T.flag[x] + fat + (long long)T.rank * T.n_blocks, epoch);
    const unsigned long long* mine = T.flag[T.rank] + fat + (long long)x * T.n_blocks;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(mine) < epoch) {
      if (global_ns() - t0 > timeout_ns) { atomicExch(T.status, 1ull); break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (y == 0 && live) {
    const volatile double* in = T.slot[T.rank] + (long long)par * T.n_ranks * T.stride + j;
    double tot = 0.0;
    for (int r = 0; r < T.n_ranks; ++r) tot += in[(long long)r * T.stride];
    if (*((volatile unsigned long long*)T.status) != 0ull) tot = __longlong_as_double(0x7ff8000000000000LL);   // NaN: the caller sees UDE_NONFINITE
    if (k == n_pm) loss_out[0] = tot;
    else if (grad_pm) grad_pm[k] = tot;
  }
  __syncthreads();
  if (x == 0 && y == 0) T.epoch[blockIdx.x] = epoch;
}


// ---- forecast chain (src/ModelTesting.jl:541-571 with the extrapolation damping of src/ProcessModels.jl:2-25) ----------
// lim = [umax | umin | umean] (3 d doubles); x is the running state
__global__ void ude_chain_set(const double* __restrict__ x, double* __restrict__ start_col, int d) {
  if (threadIdx.x < d) start_col[threadIdx.x] = x[threadIdx.x];
}
__global__ void ude_chain_damp(double* __restrict__ x, const double* __restrict__ pred, const double* __restrict__ lim, int d,
                               double l, double rho, double* __restrict__ out) {
  const int i = threadIdx.x;
  if (i >= d) return;
  const double xi = x[i], umax = lim[i], umin = lim[d + i], umean = lim[2 * d + i], span = umax - umin;
  double du = pred[i] - xi;
  if (xi > umax) { const double z = (xi - umax) / span, w = exp(-0.5 / (l * l) * z * z); du = w * du - rho * (1.0 - w) * (xi - umean); }
  if (xi < umin) { const double z = (xi - umin) / span, w = exp(-0.5 / (l * l) * z * z); du = w * du - (1.0 - w) * rho * (xi - umean); }
  const double xn = xi + du;
  x[i] = xn; out[i] = xn;
}

// Optimisers.jl Adam: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; theta -= eta * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)
__global__ void ude_adam_update(double* __restrict__ theta, double* __restrict__ m, double* __restrict__ v,
                                const double* __restrict__ g, long long n, double eta, double c1, double c2) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double gi = g[i];
    const double mi = 0.9 * m[i] + 0.1 * gi;
    const double vi = 0.999 * v[i] + 0.001 * gi * gi;
    m[i] = mi; v[i] = vi;
    theta[i] -= eta * (mi / c1) / (sqrt(vi / c2) + 1e-8);
  }
}

// Graph-replayable variants: the bias-correction state (beta1^t, beta2^t) and the iteration counter live on the device,
// so one captured iteration can be replayed n_iter times with no host-side parameter changes.
__global__ void ude_adam_update_dev(double* __restrict__ theta, double* __restrict__ m, double* __restrict__ v,
                                    const double* __restrict__ g, long long n, double eta, const double* __restrict__ state) {
  // __dmul_rn: no FMA contraction, so the correction factors are bit-identical to the host loop's b *= beta; 1 - b
  const double c1 = 1.0 - __dmul_rn(state[0], 0.9), c2 = 1.0 - __dmul_rn(state[1], 0.999);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double gi = g[i];
    const double mi = 0.9 * m[i] + 0.1 * gi;
    const double vi = 0.999 * v[i] + 0.001 * gi * gi;
    m[i] = mi; v[i] = vi;
    theta[i] -= eta * (mi / c1) / (sqrt(vi / c2) + 1e-8);
  }
}
__global__ void ude_adam_advance(double* __restrict__ state, const double* __restrict__ loss, double* __restrict__ trace, int M) {
  const long long it = (long long)state[2];
  for (int m = threadIdx.x; m < M; m += blockDim.x) trace[it * M + m] = loss[m];
  __syncthreads();
  if (threadIdx.x == 0) { state[0] *= 0.9; state[1] *= 0.999; state[2] = (double)(it + 1); }
}

template <class T>
__global__ void ude_fma_peak(T* out, int iters, T a, T b) {
  T acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = fma(acc[c], a, b);
  }
  T s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += acc[c];
  if (s == (T)123.456) out[0] = s;
}

// ---- segment table -------------------------------------------------------------------------------------------------
int build_plan(ude_handle* h) {
  const ude_desc& D = h->desc;
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data has not been called");
  if (D.n_cov > 0 && !h->have_cov) return fail(h, UDE_ERR_STATE, "model has covariates but ude_set_covariates has not been called");
  if (D.loss_kind == UDE_LOSS_DERIV_MATCH && !h->have_targets) return fail(h, UDE_ERR_STATE, "derivative matching needs ude_set_targets");
  // choose the kernel specialisation
  const KernelEntry* best_g = nullptr; const KernelEntry* best_f = nullptr;
  for (const KernelEntry& e : ude::kernel_registry()) {
    if (e.rhs_kind != D.rhs_kind || e.d != D.d || e.n_cov != D.n_cov || e.nn_in != D.nn_in || e.nn_out != D.nn_out) continue;
    const int lk = D.loss_kind == UDE_LOSS_STATE_SPACE ? 0 : (D.loss_kind == UDE_LOSS_DERIV_MATCH ? 2 : 1);
    if (e.solver != D.solver || e.lk != lk || e.G * e.HPL < D.nn_hidden) continue;
    if (D.lanes_per_segment > 0 && e.G != D.lanes_per_segment) continue;
    if (D.hidden_per_lane > 0 && e.HPL != D.hidden_per_lane) continue;
    if (e.dtype != D.dtype) continue;
    if (e.single_model_only && h->n_models != 1) continue;
    if (const char* want = getenv("UDE_KERNEL")) { if (*want && !strstr(e.name, want)) continue; }   // tuning aid: pick a specialisation by name
    if (D.weights_in_smem == 1 && e.ws_bytes != 0) continue;
    if (D.weights_in_smem == 2 && e.ws_bytes == 0) continue;
    const KernelEntry*& slot = e.grad ? best_g : best_f;
    // priority first (set per specialisation from measurements), then the smallest padded hidden width
    auto key = [](const KernelEntry* k) { return (long long)k->priority * 100000 + (long long)k->G * k->HPL; };
    if (!slot || key(&e) < key(slot)) slot = &e;
  }
  if (!best_g || !best_f) {
    char buf[320];
    snprintf(buf, sizeof buf, "no compiled sm_100a kernel for rhs_kind=%d d=%d n_cov=%d nn=%d->%d->%d solver=%d (G=%d HPL=%d); "
             "add an instantiation in csrc/ude_inst_*.cu", D.rhs_kind, D.d, D.n_cov, D.nn_in, D.nn_hidden, D.nn_out, D.solver,
             D.lanes_per_segment, D.hidden_per_lane);
    return fail(h, UDE_ERR_UNSUPPORTED, buf);
  }
  h->k_grad = best_g; h->k_fwd = best_f;
  const int SPW = best_g->spw;
  const int S = D.substeps, pre = (D.loss_kind == UDE_LOSS_MULTI_SHOOT && D.preroll_eps > 0.0) ? 1 : 0;

  std::vector<Seg> segs;
  segs.reserve((size_t)h->n_cols + 64);
  int64_t steps = 0, max_steps = 0;
  std::vector<int64_t> task_steps;
  auto pad_model = [&]() {
    while (segs.size() % SPW) { Seg z = segs.back(); z.flags |= ude::SEG_NULL; z.L = 0; segs.push_back(z); }
  };
  int64_t cur_model = -1;
  h->model_task0.assign((size_t)h->n_models + 1, 0);
  for (int64_t s = 0; s < h->n_series; ++s) {
    const int64_t m = h->model_of_series[s], c0 = h->starts[s], len = h->lengths[s];
    if (m != cur_model) { if (!segs.empty()) pad_model(); cur_model = m; h->model_task0[(size_t)m] = (int64_t)segs.size() / SPW; }
    auto push = [&](int64_t col, int64_t L, uint16_t flags) {
      Seg z; z.col = (int32_t)col; z.series = (int32_t)s; z.L = (int16_t)L; z.flags = flags; z.model = (int32_t)m;
      segs.push_back(z);
    };
    switch (D.loss_kind) {
      case UDE_LOSS_STATE_SPACE:
        if (len == 1) push(c0, 0, ude::SEG_SKIP_PROC | ude::SEG_NO_PREV);
        for (int64_t t = 1; t < len; ++t) {
          uint16_t f = (t == 1) ? ude::SEG_OBS_FIRST : 0;
          if (!h->skip.empty() && h->skip[c0 + t]) f |= ude::SEG_SKIP_PROC; else steps += S;
          push(c0 + t, 1, f);
        }
        break;
      case UDE_LOSS_MULTI_SHOOT:
        for (int64_t t = 0; t < len; t += D.pred_length) {
          const int64_t L = (len - 1 - t >= D.pred_length) ? D.pred_length : (len - 1 - t);
          push(c0 + t, L, 0);
          steps += pre + L * S;
        }
        break;
      case UDE_LOSS_SHOOT:
        if (len - 1 > 32767) return fail(h, UDE_ERR_INVALID, "shooting: series longer than 32768 points");
        push(c0, len - 1, 0);
        steps += (len - 1) * S;
        break;
      case UDE_LOSS_DERIV_MATCH:
        for (int64_t t = D.remove_ends; t < len - D.remove_ends; ++t) push(c0 + t, 0, 0);
        break;
      default: return fail(h, UDE_ERR_INVALID, "unknown loss_kind");
    }
  }
  if (segs.empty()) return fail(h, UDE_ERR_INVALID, "no segments (empty data)");
  pad_model();
  h->n_tasks = (int64_t)segs.size() / SPW;
  h->model_task0[(size_t)h->n_models] = h->n_tasks;
  h->n_segments = 0;
  for (const Seg& z : segs) if (!(z.flags & ude::SEG_NULL)) ++h->n_segments;
  for (int64_t t = 0; t < h->n_tasks; ++t) {
    int64_t mx = 0;
    for (int q = 0; q < SPW; ++q) {
      const Seg& z = segs[t * SPW + q];
      int64_t st = (D.loss_kind == UDE_LOSS_STATE_SPACE) ? S : (D.loss_kind == UDE_LOSS_DERIV_MATCH ? 0 : pre + (int64_t)z.L * S);
      mx = std::max(mx, st);
    }
    max_steps = std::max(max_steps, mx);
  }
  h->steps_per_eval = steps;
  h->max_steps_per_task = max_steps;
  UDE_CUDA_TRY(h, h->d_segs.upload(segs.data(), segs.size(), h->stream));

  // launch geometry: persistent grid, a multiple of the SM count
  const int stages = D.solver == UDE_TSIT5 ? 6 : (D.solver == UDE_EULER ? 1 : 4);
  // dynamic smem: block partial row (single model) padded to 16 B + per-warp two-step TMA staging ring of the stash
  const size_t red_doubles = (size_t)(((h->n_models == 1 ? D.n_pm + 1 : 0) + 1) / 2 * 2);
  auto smem_for = [&](const KernelEntry* k) {
    const size_t warps = (size_t)(ude::kBlockThreads / 32) * (size_t)k->warp_bytes, red = red_doubles * sizeof(double);
    if (k->red_alias) return (size_t)k->block_bytes + std::max(warps, red);      // the partial row lives in the per-warp region
    return red + (size_t)k->block_bytes + warps;
  };
  h->smem_bytes = smem_for(h->k_grad);
  h->smem_bytes_fwd = smem_for(h->k_fwd);
  if (h->smem_bytes > 40 * 1024)
    UDE_CUDA_TRY(h, cudaFuncSetAttribute(h->k_grad->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
  if (h->smem_bytes_fwd > 40 * 1024)
    UDE_CUDA_TRY(h, cudaFuncSetAttribute(h->k_fwd->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes_fwd));
  int occ_g = 0, occ_f = 0;
  UDE_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&occ_g, h->k_grad->fn, ude::kBlockThreads, h->smem_bytes, 0));
  UDE_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&occ_f, h->k_fwd->fn, ude::kBlockThreads, h->smem_bytes_fwd, 0));
  if (occ_g < 1 || occ_f < 1) return fail(h, UDE_ERR_CUDA, "kernel does not fit on an SM");
  const int wpb = ude::kBlockThreads / 32;
  const int tpb = wpb / best_g->warps_per_task;            // tasks a block works on concurrently
  const int64_t need_blocks = (h->n_tasks + tpb - 1) / tpb;
  const char* env_occ = getenv("UDE_MAX_BLOCKS_PER_SM");
  if (env_occ && atoi(env_occ) > 0) { occ_g = std::min(occ_g, atoi(env_occ)); occ_f = std::min(occ_f, atoi(env_occ)); }
  h->grid_grad = (int)std::min<int64_t>(need_blocks, (int64_t)occ_g * h->sm_count);
  h->grid_fwd = (int)std::min<int64_t>(need_blocks, (int64_t)occ_f * h->sm_count);
  h->stash_rows_per_warp = (long long)max_steps * stages * h->k_grad->rows_per_stage;
  // the stash must fit: shrink the grid (whole multiples of the SM count) if a long shooting problem would not
  size_t free_b = 0, total_b = 0;
  UDE_CUDA_TRY(h, cudaMemGetInfo(&free_b, &total_b));
  while (h->grid_grad > h->sm_count &&
         (double)h->grid_grad * wpb * h->stash_rows_per_warp * 32.0 * h->k_grad->elem_bytes > 0.5 * (double)free_b)
    h->grid_grad -= h->sm_count;
  const size_t stash_bytes = (size_t)h->grid_grad * wpb * (size_t)h->stash_rows_per_warp * 32 * (size_t)h->k_grad->elem_bytes;
  if ((double)stash_bytes > 0.8 * (double)free_b) return fail(h, UDE_ERR_CUDA, "reverse-sweep stash does not fit in device memory");
  UDE_CUDA_TRY(h, h->d_stash.reserve((stash_bytes + 7) / 8));
  // ---- chunk plan for the host-buffer path (single model, big problems): K contiguous task ranges; chunk k needs the
  // uhat columns of every series it touches on the device before it starts (up_*), and after it finishes the gradient
  // of every column no later chunk touches is final (dn_*).
  h->chunks.clear();
  {
    const int64_t NT = h->n_tasks;
    int K = 1;
    if (h->n_models == 1 && D.loss_kind != UDE_LOSS_DERIV_MATCH) {
      const char* env = getenv("UDE_HOST_CHUNKS");
      K = env ? atoi(env) : (int)std::min<int64_t>(12, NT / 6000);   // measured on C3: ~12 chunks = 95% of the resident rate
      if (K < 1) K = 1;
      if ((int64_t)K > NT) K = (int)NT;
    }
    // Every chunk is one launch of the persistent grid, whose warps stride over the chunk's tasks: a chunk that is not a
    // whole number of rounds of the grid leaves part of the GPU idle in its last round (3.5 rounds per chunk = 12 % lost).
    // So chunks are whole rounds; the remainder goes to the last one.
    const char* wenv = getenv("UDE_HOST_CHUNK_ROUNDS");
    int64_t per_chunk = 0;
    std::vector<int64_t> edge_sizes;
    if (K > 1 && !(wenv && atoi(wenv) == 0)) {
      const int64_t Wg = (int64_t)h->grid_grad * (wpb / best_g->warps_per_task);
      const int64_t rounds = std::max<int64_t>(1, (NT + (int64_t)K * Wg / 2) / ((int64_t)K * Wg));
      per_chunk = rounds * Wg;
      // the first upload and the last download are the only transfers nothing overlaps: one-round chunks at both ends
      const char* eenv = getenv("UDE_HOST_CHUNK_EDGE");
      if (!(eenv && atoi(eenv) == 0) && rounds > 1 && NT > 2 * per_chunk + 2 * Wg) {
        edge_sizes.push_back(Wg);
        int64_t rem = NT - 2 * Wg;
        while (rem > 0) { const int64_t t = std::min(per_chunk, rem); edge_sizes.push_back(t); rem -= t; }
        edge_sizes.push_back(Wg);
        K = (int)edge_sizes.size();
      } else {
        K = (int)((NT + per_chunk - 1) / per_chunk);
      }
      if (K < 1) K = 1;
    }
    auto series_end_col = [&](int64_t s) { return h->starts[s] + h->lengths[s]; };
    const char* renv = getenv("UDE_HOST_CHUNK_RAMP");
    const bool ramp = K >= 4 && renv && atoi(renv) == 1;      // tuning aid, off: measured slower on C3 (9.20 ms vs 8.99 ms per step)
    // uniform chunks by default; UDE_HOST_CHUNK_RAMP=1 makes them triangular (1, 2, 3, ..., 3, 2, 1) so that the first upload
    // and the last download -- the only transfers nothing overlaps -- are small: the under-filled edge launches cost more
    // than that saves.  Boundaries are strictly increasing (K <= NT).
    std::vector<int64_t> bounds((size_t)K + 1, 0);
    {
      long long wtot = 0, w = 0;
      for (int i = 0; i < K; ++i) wtot += ramp ? std::min(i + 1, K - i) : 1;
      for (int k = 1; k <= K; ++k) {
        w += ramp ? std::min(k, K - k + 1) : 1;
        int64_t want = NT * w / wtot;
        if (!edge_sizes.empty()) want = bounds[(size_t)k - 1] + edge_sizes[(size_t)k - 1];
        else if (per_chunk > 0) want = std::min<int64_t>(per_chunk * k, NT);
        bounds[(size_t)k] = std::max<int64_t>(want, bounds[(size_t)k - 1] + 1);
      }
      bounds[(size_t)K] = NT;
      for (int k = K - 1; k >= 1; --k) bounds[(size_t)k] = std::min(bounds[(size_t)k], bounds[(size_t)k + 1] - 1);
    }
    int64_t prev_up = 0, prev_dn = 0;
    for (int k = 0; k < K; ++k) {
      ude_handle::Chunk c;
      c.task_begin = bounds[(size_t)k]; c.task_end = bounds[(size_t)k + 1];
      const int64_t s_last = segs[(size_t)c.task_end * SPW - 1].series;          // last series this chunk touches
      c.up_lo = prev_up; c.up_hi = (k == K - 1) ? h->n_cols : series_end_col(s_last);
      // a series touched by the next chunk as well (its first task may start inside s_last) is finalised there
      int64_t dn_hi = h->n_cols;
      if (k < K - 1) {
        const int64_t s_next_first = segs[(size_t)c.task_end * SPW].series;
        dn_hi = h->starts[std::min<int64_t>(s_next_first, s_last)];
        if (s_next_first > s_last) dn_hi = series_end_col(s_last);
      }
      c.dn_lo = prev_dn; c.dn_hi = std::max(dn_hi, prev_dn);
      prev_up = c.up_hi; prev_dn = c.dn_hi;
      h->chunks.push_back(c);
    }
  }
  // ---- many-model handles (folds x ensemble members, every model its own theta block): the same three-stream overlap with
  // chunks that are runs of WHOLE models -- theta / gradient of a run are one contiguous range, nothing is shared between
  // chunks.  A warp takes a contiguous ceil(tasks / warps) run of the chunk's tasks, so a chunk boundary is the last model
  // boundary at or below a whole number of rounds of the persistent grid (a chunk of r rounds + a few tasks would cost r + 1).
  if (h->n_models > 1) {
    const int64_t NT = h->n_tasks, M = h->n_models;
    const int64_t Wg = std::max<int64_t>(1, (int64_t)h->grid_grad * (wpb / best_g->warps_per_task));
    const char* env = getenv("UDE_HOST_CHUNKS");
    const char* wenv = getenv("UDE_HOST_CHUNK_ROUNDS");
    const bool whole_rounds = !(wenv && atoi(wenv) == 0);
    const int64_t rounds_total = (NT + Wg - 1) / Wg;
    int64_t K = env ? atoi(env) : std::min<int64_t>(6, rounds_total / 4);      // >= 4 rounds per chunk: small problems stay whole
    K = std::max<int64_t>(1, std::min<int64_t>(K, M));
    if (K > 1) {
      std::vector<int64_t> mb(1, 0);                                            // model boundaries of the chunks
      // the first upload and the last download are the only transfers nothing overlaps: one-round chunks at both ends
      const char* eenv = getenv("UDE_HOST_CHUNK_EDGE");
      const bool edge = whole_rounds && !(eenv && atoi(eenv) == 0) && rounds_total >= 3 * K && M >= K + 2;
      const int64_t KK = edge ? K + 2 : K;
      for (int64_t k = 1; k < KK; ++k) {
        const int64_t t_prev = h->model_task0[(size_t)mb.back()];
        int64_t left = KK - k + 1;                                              // chunks still to cut, this one included
        int64_t rounds_left = (NT - t_prev + Wg - 1) / Wg;
        int64_t want = 1;
        if (!(edge && k == 1)) {
          if (edge) { left -= 1; rounds_left -= 1; }                            // the one-round chunk at the far end
          want = std::max<int64_t>(1, (rounds_left + left / 2) / left);
        }
        const int64_t target = whole_rounds ? t_prev + want * Wg : NT * k / K;
        // last model boundary with model_task0[m] <= target
        int64_t m = (int64_t)(std::upper_bound(h->model_task0.begin(), h->model_task0.end(), target) - h->model_task0.begin()) - 1;
        m = std::min<int64_t>(std::max<int64_t>(m, mb.back() + 1), M - (KK - k));
        if (m > mb.back() && m < M) mb.push_back(m);
      }
      mb.push_back(M);
      auto th_end = [&](int64_t m) { return m < M ? h->theta_base[(size_t)m] : h->n_theta; };
      if (mb.size() > 2) h->chunks.clear();                                     // replaces the single whole-problem chunk
      if (mb.size() > 2)
        for (size_t k = 0; k + 1 < mb.size(); ++k) {
          ude_handle::Chunk c;
          c.m_lo = mb[k]; c.m_hi = mb[k + 1];
          c.task_begin = h->model_task0[(size_t)c.m_lo]; c.task_end = h->model_task0[(size_t)c.m_hi];
          c.up_lo = c.dn_lo = th_end(c.m_lo); c.up_hi = c.dn_hi = th_end(c.m_hi);
          h->chunks.push_back(c);
        }
    }
  }
  UDE_CUDA_TRY(h, h->d_block_part.reserve((size_t)std::max(h->grid_grad, h->grid_fwd) * (size_t)(D.n_pm + 1) * std::max<size_t>(1, h->chunks.size())));
  if (h->chunks.size() > 1 && !h->s_h2d) {
    UDE_CUDA_TRY(h, cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    UDE_CUDA_TRY(h, cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    UDE_CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
  }
  while (h->ev_up.size() < h->chunks.size()) {
    cudaEvent_t a, b;
    UDE_CUDA_TRY(h, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    UDE_CUDA_TRY(h, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    h->ev_up.push_back(a); h->ev_done.push_back(b);
  }
  UDE_CUDA_TRY(h, h->d_model_loss.reserve((size_t)h->n_models));
  UDE_CUDA_TRY(h, h->d_red.reserve((size_t)D.n_pm + 2));
  if ((h->comm || h->x_attached) && D.has_series_param && D.off_series_param + h->n_series != D.n_pm)
    return fail(h, UDE_ERR_INVALID, "multi-GPU: the per-series parameter vector must be the last entry of process_model");
  UDE_CUDA_TRY(h, cudaStreamSynchronize(h->stream));   // segs came from a local vector
  h->dirty = false;
  return UDE_OK;
}

DevProblem make_dev_problem(const ude_handle* h, const double* d_theta, double* d_grad, double* d_traj) {
  const ude_desc& D = h->desc;
  DevProblem P{};
  P.d = D.d; P.n_cov = D.n_cov; P.nn_in = D.nn_in; P.H = D.nn_hidden; P.nn_out = D.nn_out; P.n_scalars = D.n_scalars;
  P.has_series_param = D.has_series_param; P.substeps = D.substeps;
  P.off_w1 = D.off_w1; P.off_b1 = D.off_b1; P.off_w2 = D.off_w2; P.off_b2 = D.off_b2;
  for (int k = 0; k < 8; ++k) P.off_scalar[k] = D.off_scalar[k];
  P.off_series_param = D.off_series_param; P.n_pm = D.n_pm;
  P.time_scale = D.time_scale; P.time_shift = D.time_shift;
  P.loss_kind = D.loss_kind; P.pred_length = D.pred_length; P.proc_inv_dt = D.proc_inv_dt;
  P.shoot_from_theta = D.shoot_from_theta; P.shoot_first_scored = D.shoot_first_scored; P.remove_ends = D.remove_ends;
  P.preroll_eps = D.preroll_eps;
  const int d = D.d;
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      P.A[a + d * b] = D.A[a + d * b]; P.Asym[a + d * b] = D.A[a + d * b] + D.A[b + d * a];
      P.B[a + d * b] = D.B[a + d * b]; P.Bsym[a + d * b] = D.B[a + d * b] + D.B[b + d * a];
    }
  P.starts = h->d_starts.p; P.lengths = h->d_lengths.p; P.theta_base = h->d_theta_base.p;
  P.model_col0 = h->d_model_col0.p; P.model_series0 = h->d_model_series0.p;
  P.times = h->d_times.p; P.data = h->d_data.p;
  P.obs_scale = h->d_obs_scale.p; P.proc_scale = h->d_proc_scale.p; P.shoot_weight = h->d_shoot_weight.p;
  P.knot_off = h->d_knot_off.p; P.knot_t = h->d_knot_t.p; P.knot_x = h->d_knot_x.p; P.knot_inv = h->d_knot_inv.p;
  P.dm_u = h->d_dm_u.p; P.dm_dudt = h->d_dm_dudt.p;
  P.segs = h->d_segs.p; P.task_begin = 0; P.n_tasks = h->n_tasks; P.n_models = (int)h->n_models;
  P.n_cols_total = h->n_cols;
  P.theta = d_theta; P.grad = d_grad; P.traj_out = d_traj;
  P.uhat_x = nullptr; P.guhat_x = nullptr;
  P.block_part = h->d_block_part.p; P.model_loss = h->d_model_loss.p;
  P.stash = h->d_stash.p; P.stash_rows_per_warp = h->stash_rows_per_warp;
  { const char* e = getenv("UDE_STASH_DISCARD"); P.stash_discard = e ? atoi(e) : 1; }
  return P;
}

int finalize_eval(ude_handle* h, const double* d_theta, double* d_loss, double* d_grad, int n_rows, cudaStream_t s);

// launch the segment kernel over tasks [t0, t1) writing its per-block rows at row offset `row0`
int launch_segments(ude_handle* h, const double* d_theta, double* d_grad, double* d_traj, int64_t t0, int64_t t1, int row0,
                    cudaStream_t s, const double* uhat_x = nullptr, double* guhat_x = nullptr) {
  const bool want_grad = d_grad != nullptr;
  const KernelEntry* k = want_grad ? h->k_grad : h->k_fwd;
  const int grid = want_grad ? h->grid_grad : h->grid_fwd;
  DevProblem P = make_dev_problem(h, d_theta, d_grad, d_traj);
  P.task_begin = t0; P.n_tasks = t1;
  P.uhat_x = uhat_x; P.guhat_x = guhat_x;
  P.block_part = h->d_block_part.p + (size_t)row0 * (size_t)(h->desc.n_pm + 1);
  void* args[] = {&P};
  const bool timed = h->ev_used < h->ev_cap;
  if (timed) UDE_CUDA_TRY(h, cudaEventRecord(h->ev0[h->ev_used], s));
  UDE_CUDA_TRY(h, cudaLaunchKernel(k->fn, dim3(grid), dim3(ude::kBlockThreads), args, want_grad ? h->smem_bytes : h->smem_bytes_fwd, s));
  if (timed) { UDE_CUDA_TRY(h, cudaEventRecord(h->ev1[h->ev_used], s)); ++h->ev_used; }
  return UDE_OK;
}

// enqueue one loss(+grad) evaluation on `s`; all pointers are device pointers
int enqueue_eval(ude_handle* h, const double* d_theta, double* d_loss, double* d_grad, double* d_traj, cudaStream_t s) {
  if (h->dirty) { int rc = build_plan(h); if (rc != UDE_OK) return rc; }
  const bool want_grad = d_grad != nullptr;
  const int grid = want_grad ? h->grid_grad : h->grid_fwd;
  if (want_grad) UDE_CUDA_TRY(h, cudaMemsetAsync(d_grad, 0, (size_t)h->n_theta * sizeof(double), s));
  if (h->n_models > 1) UDE_CUDA_TRY(h, cudaMemsetAsync(h->d_model_loss.p, 0, (size_t)h->n_models * sizeof(double), s));
  int rc = launch_segments(h, d_theta, d_grad, d_traj, 0, h->n_tasks, 0, s);
  if (rc != UDE_OK) return rc;
  // trajectories only (ude_forward): the loss is not used, so no reduction and -- on a peer-attached / NCCL handle -- no
  // cross-rank exchange: a forward call is NOT a collective and may be made by one rank alone
  if (d_traj != nullptr && !want_grad) return UDE_OK;
  return finalize_eval(h, d_theta, d_loss, d_grad, grid, s);
}

int finalize_eval(ude_handle* h, const double* d_theta, double* d_loss, double* d_grad, int grid, cudaStream_t s) {
  const ude_desc& D = h->desc;
  const bool want_grad = d_grad != nullptr;
  const long long n_w1 = (long long)D.nn_hidden * D.nn_in, n_w2 = (long long)D.nn_out * D.nn_hidden;
  if (h->n_models == 1) {
    const long long pmb = (long long)D.d * h->n_cols;
    const long long slo = D.has_series_param ? D.off_series_param : -1, shi = D.has_series_param ? slo + h->n_series : -1;
    const bool use_comm = h->comm != nullptr && !h->x_attached;
    // with a communicator the reduce buffer is [loss | shared process-model gradient]: the per-series vector (which
    // must be the tail of process_model) is rank-local and stays out of the exchange
    const long long n_shared = D.has_series_param ? D.off_series_param : D.n_pm;
    double* gp = use_comm ? h->d_red.p + 1 : (want_grad ? d_grad + pmb : h->d_red.p + 1);
    double* lp = use_comm ? h->d_red.p : d_loss;
    const int nb = (int)((D.n_pm + 1 + 31) / 32);
    if (h->x_attached) {
      PeerTable T;
      const size_t slots = 2ull * h->x_ranks * (size_t)h->x_stride;                 // doubles
      const size_t flags = 2ull * h->x_ranks * (size_t)h->x_blocks;                 // u64
      for (int r = 0; r < UDE_MAX_PEERS; ++r) {
        T.slot[r] = r < h->x_ranks ? (double*)h->x_peer[r] : nullptr;
        T.flag[r] = r < h->x_ranks ? (unsigned long long*)h->x_peer[r] + slots : nullptr;
      }
      T.epoch = (unsigned long long*)h->x_local + slots + flags;
      T.status = T.epoch + h->x_blocks;
      T.stride = h->x_stride; T.n_ranks = h->x_ranks; T.rank = h->rank; T.n_blocks = h->x_blocks;
      const char* te = getenv("UDE_PEER_TIMEOUT_MS");
      const unsigned long long tmo = (unsigned long long)(te ? atoll(te) : 60000) * 1000000ull;
      ude_finalize_exchange<<<h->x_blocks, dim3(32, 8), 0, s>>>(h->d_block_part.p, grid, D.n_pm, n_shared, d_theta + pmb, D.off_w1, n_w1,
                                                                D.off_w2, n_w2, h->d_reg_weight.p, want_grad ? d_grad + pmb : nullptr,
                                                                d_loss, T, tmo);
      UDE_CUDA_TRY(h, cudaGetLastError());
      return UDE_OK;
    }
    // rank r > 0 must not add the regulariser again: only rank 0 carries it (host passes reg_weight = 0 elsewhere)
    ude_finalize_single<<<nb, dim3(32, 8), 0, s>>>(h->d_block_part.p, grid, D.n_pm, d_theta + pmb, D.off_w1, n_w1, D.off_w2, n_w2,
                                                   slo, shi, h->d_reg_weight.p, gp, lp, 0);
    UDE_CUDA_TRY(h, cudaGetLastError());
    if (use_comm) {
      int rc = g_nccl.AllReduce(h->d_red.p, h->d_red.p, (size_t)n_shared + 1, /*ncclFloat64*/ 8, /*ncclSum*/ 0, h->comm, s);
      if (rc != 0) return fail(h, UDE_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
      const int nb2 = (int)((D.n_pm + 1 + 255) / 256);
      ude_scatter_reduced<<<nb2, 256, 0, s>>>(h->d_red.p, D.n_pm, slo, shi, want_grad ? d_grad + pmb : nullptr, d_loss);
      UDE_CUDA_TRY(h, cudaGetLastError());
    }
  } else {
    ude_finalize_multi<<<(int)h->n_models, 128, 0, s>>>(d_theta, d_grad, h->d_model_loss.p, h->d_theta_base.p, h->d_model_col0.p, D.d,
                                                        D.off_w1, n_w1, D.off_w2, n_w2, h->d_reg_weight.p, d_loss);
    UDE_CUDA_TRY(h, cudaGetLastError());
  }
  return UDE_OK;
}

}  // namespace

#include "ude_ukf.cuh"

// other translation units of the library (ude_smooth.cu) report handle-less failures through the same channel
int ude_fail_create(int code, const std::string& msg) { return fail(nullptr, code, msg); }

// ==================================================================================================================
#pragma GCC visibility push(default)
extern "C" {

int ude_abi_version(void) { return UDE_ABI_VERSION; }

int ude_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* ude_last_error(const ude_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ude_create(const ude_desc* desc, ude_handle** out) {
  if (!desc || !out) return fail(nullptr, UDE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (desc->abi_version != UDE_ABI_VERSION || desc->struct_bytes != (int32_t)sizeof(ude_desc))
    return fail(nullptr, UDE_ERR_INVALID, "ude_desc ABI mismatch (abi_version / struct_bytes)");
  if (desc->d < 1 || desc->d > UDE_MAX_D || desc->nn_out < 1 || desc->nn_out > UDE_MAX_D || desc->nn_hidden < 1 ||
      desc->substeps < 1 || desc->n_scalars < 0 || desc->n_scalars > UDE_MAX_SCALARS || desc->n_cov < 0)
    return fail(nullptr, UDE_ERR_INVALID, "bad model dimensions");
  if (desc->dtype != UDE_F64 && desc->dtype != UDE_F32) return fail(nullptr, UDE_ERR_INVALID, "bad dtype");
  if (desc->solver != UDE_RK4 && desc->solver != UDE_TSIT5 && desc->solver != UDE_EULER) return fail(nullptr, UDE_ERR_INVALID, "bad solver");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, UDE_ERR_NO_DEVICE, std::string("no CUDA device (libude_cuda has no CPU fallback): ") + cudaGetErrorString(e));
  }
  if (desc->device < 0 || desc->device >= n) return fail(nullptr, UDE_ERR_INVALID, "device ordinal out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, desc->device)) != cudaSuccess) return fail(nullptr, UDE_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10) {
    char buf[160];
    snprintf(buf, sizeof buf, "device %d is sm_%d%d; libude_cuda is built for sm_100a only", desc->device, prop.major, prop.minor);
    return fail(nullptr, UDE_ERR_NO_DEVICE, buf);
  }
  ude_handle* h = new ude_handle();
  h->desc = *desc; h->device = desc->device; h->sm_count = prop.multiProcessorCount;
  if ((e = cudaSetDevice(h->device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    std::string m = cudaGetErrorString(e); delete h; return fail(nullptr, UDE_ERR_CUDA, m);
  }
  *out = h;
  return UDE_OK;
}

int ude_destroy(ude_handle* h) {
  if (!h) return UDE_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);      // nothing of this handle may still be running when its buffers go
  if (h->s_h2d) cudaStreamSynchronize(h->s_h2d);
  if (h->s_d2h) cudaStreamSynchronize(h->s_d2h);
  delete h->ukf;
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (int r = 0; r < UDE_MAX_PEERS; ++r) if (h->x_opened[r] && h->x_peer[r]) cudaIpcCloseMemHandle(h->x_peer[r]);
  if (h->x_local) cudaFree(h->x_local);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  for (cudaEvent_t e : h->ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_done) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev0) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev1) cudaEventDestroy(e);
  delete h;
  return UDE_OK;
}

int ude_set_data(ude_handle* h, int64_t n_series, const int64_t* starts, const int64_t* lengths, const int64_t* model_of_series,
                 int64_t n_cols, const double* times, const double* data) {
  if (!h) return UDE_ERR_INVALID;
  if (n_series < 1 || n_cols < 1 || !starts || !lengths || !times || !data) return fail(h, UDE_ERR_INVALID, "empty or null data");
  if (n_cols > 2000000000LL) return fail(h, UDE_ERR_INVALID, "more than 2^31 columns per handle: shard the series");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  const int d = h->desc.d;
  h->starts.assign(starts, starts + n_series);
  h->lengths.assign(lengths, lengths + n_series);
  h->model_of_series.assign(n_series, 0);
  if (model_of_series) h->model_of_series.assign(model_of_series, model_of_series + n_series);
  int64_t expect = 0;
  for (int64_t s = 0; s < n_series; ++s) {
    if (starts[s] != expect || lengths[s] < 1) return fail(h, UDE_ERR_INVALID, "series must tile the columns contiguously, each with >= 1 point");
    expect += lengths[s];
    const int64_t m = h->model_of_series[s], prev = s ? h->model_of_series[s - 1] : 0;
    if ((s == 0 && m != 0) || m < prev || m > prev + 1) return fail(h, UDE_ERR_INVALID, "model_of_series must be sorted, dense, 0-based");
  }
  if (expect != n_cols) return fail(h, UDE_ERR_INVALID, "sum(lengths) != n_cols");
  const int64_t M = h->model_of_series.back() + 1;
  h->n_series = n_series; h->n_cols = n_cols; h->n_models = M;
  h->model_col0.assign(M + 1, 0); h->model_series0.assign(M + 1, 0); h->theta_base.assign(M, 0);
  for (int64_t s = n_series - 1; s >= 0; --s) { h->model_col0[h->model_of_series[s]] = starts[s]; h->model_series0[h->model_of_series[s]] = s; }
  h->model_col0[M] = n_cols; h->model_series0[M] = n_series;
  int64_t off = 0;
  for (int64_t m = 0; m < M; ++m) { h->theta_base[m] = off; off += (int64_t)d * (h->model_col0[m + 1] - h->model_col0[m]) + h->desc.n_pm; }
  h->n_theta = off;
  if (h->desc.has_series_param) {
    for (int64_t m = 0; m < M; ++m)
      if (h->desc.off_series_param + (h->model_series0[m + 1] - h->model_series0[m]) > h->desc.n_pm)
        return fail(h, UDE_ERR_INVALID, "per-series parameter vector does not fit inside process_model");
  }
  cudaStream_t s = h->stream;
  UDE_CUDA_TRY(h, h->d_starts.upload((const long long*)h->starts.data(), n_series, s));
  UDE_CUDA_TRY(h, h->d_lengths.upload((const long long*)h->lengths.data(), n_series, s));
  UDE_CUDA_TRY(h, h->d_theta_base.upload((const long long*)h->theta_base.data(), M, s));
  UDE_CUDA_TRY(h, h->d_model_col0.upload((const long long*)h->model_col0.data(), M + 1, s));
  UDE_CUDA_TRY(h, h->d_model_series0.upload((const long long*)h->model_series0.data(), M + 1, s));
  UDE_CUDA_TRY(h, h->d_times.upload(times, n_cols, s));
  UDE_CUDA_TRY(h, h->d_data.upload(data, (size_t)d * n_cols, s));
  UDE_CUDA_TRY(h, h->d_theta.reserve((size_t)h->n_theta));
  UDE_CUDA_TRY(h, h->d_grad.reserve((size_t)h->n_theta));
  UDE_CUDA_TRY(h, h->d_loss.reserve((size_t)M));
  h->have_data = true; h->dirty = true;
  h->skip.clear();
  ukf_invalidate(h);
  int rc = ude_set_model_weights(h, nullptr, nullptr, nullptr, nullptr);
  if (rc != UDE_OK) return rc;
  UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
  return UDE_OK;
}

int ude_set_model_weights(ude_handle* h, const double* obs_scale, const double* proc_scale, const double* shoot_weight,
                          const double* reg_weight) {
  if (!h) return UDE_ERR_INVALID;
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data first");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  const size_t M = (size_t)h->n_models;
  auto fill = [&](std::vector<double>& v, const double* src, double dflt) { v.assign(M, dflt); if (src) v.assign(src, src + M); };
  fill(h->obs_scale, obs_scale, 1.0); fill(h->proc_scale, proc_scale, 1.0);
  fill(h->shoot_weight, shoot_weight, 1.0); fill(h->reg_weight, reg_weight, 0.0);
  UDE_CUDA_TRY(h, h->d_obs_scale.upload(h->obs_scale.data(), M, h->stream));
  UDE_CUDA_TRY(h, h->d_proc_scale.upload(h->proc_scale.data(), M, h->stream));
  UDE_CUDA_TRY(h, h->d_shoot_weight.upload(h->shoot_weight.data(), M, h->stream));
  UDE_CUDA_TRY(h, h->d_reg_weight.upload(h->reg_weight.data(), M, h->stream));
  UDE_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return UDE_OK;
}

int ude_set_covariates(ude_handle* h, const int64_t* knot_off, const double* knot_t, const double* knot_x) {
  if (!h) return UDE_ERR_INVALID;
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data first");
  if (h->desc.n_cov < 1) return fail(h, UDE_ERR_INVALID, "model has no covariates");
  if (!knot_off || !knot_t || !knot_x) return fail(h, UDE_ERR_INVALID, "null covariate arrays");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  const size_t n_off = (size_t)h->n_series * h->desc.n_cov + 1;
  const size_t nk = (size_t)knot_off[n_off - 1];
  std::vector<double> inv(nk ? nk : 1, 0.0);
  for (size_t i = 0; i + 1 < n_off; ++i) {
    if (knot_off[i + 1] < knot_off[i]) return fail(h, UDE_ERR_INVALID, "knot_off must be non-decreasing");
    for (int64_t k = knot_off[i]; k + 1 < knot_off[i + 1]; ++k) {
      const double dt = knot_t[k + 1] - knot_t[k];
      if (!(dt > 0.0)) return fail(h, UDE_ERR_INVALID, "covariate knot times must be strictly increasing");
      inv[k] = 1.0 / dt;
    }
  }
  UDE_CUDA_TRY(h, h->d_knot_off.upload((const long long*)knot_off, n_off, h->stream));
  UDE_CUDA_TRY(h, h->d_knot_t.upload(knot_t, nk, h->stream));
  UDE_CUDA_TRY(h, h->d_knot_x.upload(knot_x, nk, h->stream));
  UDE_CUDA_TRY(h, h->d_knot_inv.upload(inv.data(), nk, h->stream));
  UDE_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->have_cov = true;
  ukf_invalidate(h);
  return UDE_OK;
}

int ude_set_targets(ude_handle* h, const double* smoothed_states, const double* dudt) {
  if (!h) return UDE_ERR_INVALID;
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data first");
  if (!smoothed_states || !dudt) return fail(h, UDE_ERR_INVALID, "null targets");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  const size_t n = (size_t)h->desc.d * h->n_cols;
  UDE_CUDA_TRY(h, h->d_dm_u.upload(smoothed_states, n, h->stream));
  UDE_CUDA_TRY(h, h->d_dm_dudt.upload(dudt, n, h->stream));
  UDE_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->have_targets = true;
  return UDE_OK;
}

int ude_set_skip(ude_handle* h, const uint8_t* skip_mask) {
  if (!h) return UDE_ERR_INVALID;
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data first");
  if (skip_mask) h->skip.assign(skip_mask, skip_mask + h->n_cols); else h->skip.clear();
  h->dirty = true;
  return UDE_OK;
}

int ude_set_option(ude_handle* h, int32_t option, int64_t value) {
  if (!h) return UDE_ERR_INVALID;
  switch (option) {
    case UDE_OPT_SPARSE_UHAT_IO: h->sparse_uhat_io = value != 0; return UDE_OK;
    case UDE_OPT_UKF_CAUSAL_INTERVAL:
      if (h->ukf_causal != (value != 0)) { h->ukf_causal = value != 0; ukf_invalidate(h); }
      return UDE_OK;
    default: return fail(h, UDE_ERR_INVALID, "unknown option");
  }
}

int ude_last_io_bytes(const ude_handle* h, int64_t* h2d_bytes, int64_t* d2h_bytes) {
  if (!h) return UDE_ERR_INVALID;
  if (h2d_bytes) *h2d_bytes = h->last_h2d;
  if (d2h_bytes) *d2h_bytes = h->last_d2h;
  return UDE_OK;
}

int64_t ude_n_theta(const ude_handle* h) { return h ? h->n_theta : 0; }
int64_t ude_n_models(const ude_handle* h) { return h ? h->n_models : 0; }
int64_t ude_steps_per_eval(const ude_handle* h) {
  if (!h) return 0;
  if (h->dirty && build_plan(const_cast<ude_handle*>(h)) != UDE_OK) return -1;
  return h->steps_per_eval;
}
int32_t ude_launches_per_eval(const ude_handle* h) {
  if (!h) return 0;
  // segment kernel + finalize (+ scatter after the NCCL allreduce; the fused peer exchange needs no extra launch)
  return (h->n_models == 1) ? ((h->comm && !h->x_attached) ? 3 : 2) : 2;
}

int ude_loss_grad_device(ude_handle* h, const double* d_theta, double* d_loss, double* d_grad, void* stream) {
  if (!h) return UDE_ERR_INVALID;
  if (!d_theta || !d_loss) return fail(h, UDE_ERR_INVALID, "null device pointer");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  return enqueue_eval(h, d_theta, d_loss, d_grad, nullptr, stream ? (cudaStream_t)stream : h->stream);
}

int ude_loss_grad(ude_handle* h, const double* theta, int64_t n_theta, double* loss, double* grad) {
  if (!h) return UDE_ERR_INVALID;
  if (!theta || !loss) return fail(h, UDE_ERR_INVALID, "null argument");
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data has not been called");
  if (n_theta != h->n_theta) {
    char buf[128]; snprintf(buf, sizeof buf, "n_theta = %lld, expected %lld", (long long)n_theta, (long long)h->n_theta);
    return fail(h, UDE_ERR_INVALID, buf);
  }
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  if (h->dirty) { int rc = build_plan(h); if (rc != UDE_OK) return rc; }
  cudaStream_t s = h->stream;
  const size_t K = h->chunks.size();
  // ---- zero-copy path (UDE_OPT_SPARSE_UHAT_IO): the kernels read the block-start columns of uhat from, and write their
  // gradient to, the caller's page-locked buffers; only the process model is copied.  One launch, no chunk pipeline.
  const double* z_theta = nullptr; double* z_grad = nullptr;
  if (grad && h->sparse_uhat_io && h->n_models == 1 &&
      (h->desc.loss_kind == UDE_LOSS_MULTI_SHOOT || h->desc.loss_kind == UDE_LOSS_SHOOT)) {
    // queried on every call: a cached answer would go stale if the caller frees a pinned buffer and reuses the address
    cudaPointerAttributes at{}, ag{};
    if (cudaPointerGetAttributes(&at, theta) == cudaSuccess && cudaPointerGetAttributes(&ag, grad) == cudaSuccess &&
        at.type == cudaMemoryTypeHost && ag.type == cudaMemoryTypeHost && at.devicePointer && ag.devicePointer) {
      z_theta = (const double*)at.devicePointer; z_grad = (double*)ag.devicePointer;
    }
    cudaGetLastError();      // a pageable pointer makes cudaPointerGetAttributes report an error on old drivers: not ours
  }
  if (z_theta) {
    const int d = h->desc.d;
    const size_t pmb = (size_t)d * h->n_cols, npm = (size_t)h->desc.n_pm;
    double* dth = h->d_theta.p; double* dg = h->d_grad.p;
    if (h->h_stage_n < npm + 1) {
      if (h->h_stage) cudaFreeHost(h->h_stage);
      h->h_stage = nullptr; h->h_stage_n = 0;
      UDE_CUDA_TRY(h, cudaMallocHost((void**)&h->h_stage, (npm + 1) * sizeof(double)));
      h->h_stage_n = npm + 1;
    }
    UDE_CUDA_TRY(h, cudaMemcpyAsync(dth + pmb, theta + pmb, npm * sizeof(double), cudaMemcpyHostToDevice, s));
    if (h->desc.has_series_param)                                              // the per-series tail is accumulated by atomics
      UDE_CUDA_TRY(h, cudaMemsetAsync(dg + pmb, 0, npm * sizeof(double), s));
    int rc = launch_segments(h, dth, dg, nullptr, 0, h->n_tasks, 0, s, z_theta, z_grad);
    if (rc != UDE_OK) return rc;
    if (!h->desc.has_series_param) {
      // finalize writes [loss | process-model gradient] into ONE contiguous device buffer (d_red): the gradient pointer is
      // biased so that finalize's `grad + pmb` lands on d_red + 1; one device-to-host copy into pinned staging per call
      rc = finalize_eval(h, dth, h->d_red.p, h->d_red.p + 1 - (ptrdiff_t)pmb, h->grid_grad, s);
      if (rc != UDE_OK) return rc;
      UDE_CUDA_TRY(h, cudaMemcpyAsync(h->h_stage, h->d_red.p, (npm + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
      UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
      loss[0] = h->h_stage[0];
      memcpy(grad + pmb, h->h_stage + 1, npm * sizeof(double));
    } else {
      rc = finalize_eval(h, dth, h->d_loss.p, dg, h->grid_grad, s);
      if (rc != UDE_OK) return rc;
      UDE_CUDA_TRY(h, cudaMemcpyAsync(loss, h->d_loss.p, sizeof(double), cudaMemcpyDeviceToHost, s));
      UDE_CUDA_TRY(h, cudaMemcpyAsync(grad + pmb, dg + pmb, npm * sizeof(double), cudaMemcpyDeviceToHost, s));
      UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
    }
    const bool reads_uhat = h->desc.loss_kind == UDE_LOSS_MULTI_SHOOT || h->desc.shoot_from_theta;
    const int64_t cols = reads_uhat ? h->n_segments : 0;
    h->last_h2d = 8 * ((int64_t)d * cols + (int64_t)npm);
    h->last_d2h = 8 * ((int64_t)d * cols + (int64_t)npm + 1);
  } else
  if (K > 1 && grad && h->n_models > 1) {
    // ---- pipelined host path, many-model handle: chunk = run of whole models; theta of run k+1 goes up and the gradient
    // and losses of run k-1 come down while run k is on the SMs.  Each run is finalised (regulariser) right behind its kernel.
    const ude_desc& D = h->desc;
    const long long n_w1 = (long long)D.nn_hidden * D.nn_in, n_w2 = (long long)D.nn_out * D.nn_hidden;
    double* dth = h->d_theta.p; double* dg = h->d_grad.p;
    UDE_CUDA_TRY(h, cudaEventRecord(h->ev_start, s));
    UDE_CUDA_TRY(h, cudaStreamWaitEvent(h->s_h2d, h->ev_start, 0));
    UDE_CUDA_TRY(h, cudaStreamWaitEvent(h->s_d2h, h->ev_start, 0));
    UDE_CUDA_TRY(h, cudaMemsetAsync(dg, 0, (size_t)n_theta * sizeof(double), s));
    UDE_CUDA_TRY(h, cudaMemsetAsync(h->d_model_loss.p, 0, (size_t)h->n_models * sizeof(double), s));
    for (size_t k = 0; k < K; ++k) {
      const ude_handle::Chunk& c = h->chunks[k];
      const size_t nb = (size_t)(c.up_hi - c.up_lo) * sizeof(double);
      UDE_CUDA_TRY(h, cudaMemcpyAsync(dth + c.up_lo, theta + c.up_lo, nb, cudaMemcpyHostToDevice, h->s_h2d));
      UDE_CUDA_TRY(h, cudaEventRecord(h->ev_up[k], h->s_h2d));
      UDE_CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_up[k], 0));
      int rc = launch_segments(h, dth, dg, nullptr, c.task_begin, c.task_end, 0, s);
      if (rc != UDE_OK) return rc;
      ude_finalize_multi<<<(int)(c.m_hi - c.m_lo), 128, 0, s>>>(dth, dg, h->d_model_loss.p + c.m_lo, h->d_theta_base.p + c.m_lo,
                                                               h->d_model_col0.p + c.m_lo, D.d, D.off_w1, n_w1, D.off_w2, n_w2,
                                                               h->d_reg_weight.p ? h->d_reg_weight.p + c.m_lo : nullptr, h->d_loss.p + c.m_lo);
      UDE_CUDA_TRY(h, cudaGetLastError());
      UDE_CUDA_TRY(h, cudaEventRecord(h->ev_done[k], s));
      UDE_CUDA_TRY(h, cudaStreamWaitEvent(h->s_d2h, h->ev_done[k], 0));
      UDE_CUDA_TRY(h, cudaMemcpyAsync(grad + c.dn_lo, dg + c.dn_lo, nb, cudaMemcpyDeviceToHost, h->s_d2h));
      UDE_CUDA_TRY(h, cudaMemcpyAsync(loss + c.m_lo, h->d_loss.p + c.m_lo, (size_t)(c.m_hi - c.m_lo) * sizeof(double), cudaMemcpyDeviceToHost, h->s_d2h));
    }
    UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
    UDE_CUDA_TRY(h, cudaStreamSynchronize(h->s_d2h));
    h->last_h2d = 8 * n_theta; h->last_d2h = 8 * (n_theta + h->n_models);
  } else
  if (K > 1 && grad) {
    // ---- pipelined host path: chunk k+1's uhat columns are uploaded and chunk k-1's gradient columns downloaded while
    // chunk k is on the SMs (PCIe is full duplex; the copies use their own streams and the copy engines).
    const int d = h->desc.d;
    const size_t pmb = (size_t)d * h->n_cols;                    // single model: theta = [uhat | process_model]
    double* dth = h->d_theta.p; double* dg = h->d_grad.p;
    UDE_CUDA_TRY(h, cudaEventRecord(h->ev_start, s));            // orders this call after earlier work on the handle's stream
    UDE_CUDA_TRY(h, cudaStreamWaitEvent(h->s_h2d, h->ev_start, 0));
    UDE_CUDA_TRY(h, cudaStreamWaitEvent(h->s_d2h, h->ev_start, 0));
    UDE_CUDA_TRY(h, cudaMemcpyAsync(dth + pmb, theta + pmb, (size_t)h->desc.n_pm * sizeof(double), cudaMemcpyHostToDevice, h->s_h2d));
    UDE_CUDA_TRY(h, cudaMemsetAsync(dg, 0, (size_t)n_theta * sizeof(double), s));
    for (size_t k = 0; k < K; ++k) {
      const ude_handle::Chunk& c = h->chunks[k];
      if (c.up_hi > c.up_lo)
        UDE_CUDA_TRY(h, cudaMemcpyAsync(dth + (size_t)d * c.up_lo, theta + (size_t)d * c.up_lo, (size_t)d * (c.up_hi - c.up_lo) * sizeof(double),
                                        cudaMemcpyHostToDevice, h->s_h2d));
      UDE_CUDA_TRY(h, cudaEventRecord(h->ev_up[k], h->s_h2d));
      UDE_CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_up[k], 0));
      int rc = launch_segments(h, dth, dg, nullptr, c.task_begin, c.task_end, (int)k * h->grid_grad, s);
      if (rc != UDE_OK) return rc;
      UDE_CUDA_TRY(h, cudaEventRecord(h->ev_done[k], s));
      if (c.dn_hi > c.dn_lo) {
        UDE_CUDA_TRY(h, cudaStreamWaitEvent(h->s_d2h, h->ev_done[k], 0));
        UDE_CUDA_TRY(h, cudaMemcpyAsync(grad + (size_t)d * c.dn_lo, dg + (size_t)d * c.dn_lo, (size_t)d * (c.dn_hi - c.dn_lo) * sizeof(double),
                                        cudaMemcpyDeviceToHost, h->s_d2h));
      }
    }
    int rc = finalize_eval(h, dth, h->d_loss.p, dg, (int)K * h->grid_grad, s);
    if (rc != UDE_OK) return rc;
    UDE_CUDA_TRY(h, cudaMemcpyAsync(loss, h->d_loss.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    UDE_CUDA_TRY(h, cudaMemcpyAsync(grad + pmb, dg + pmb, (size_t)h->desc.n_pm * sizeof(double), cudaMemcpyDeviceToHost, s));
    UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
    UDE_CUDA_TRY(h, cudaStreamSynchronize(h->s_d2h));
    h->last_h2d = 8 * n_theta; h->last_d2h = 8 * (n_theta + 1);
  } else {
    UDE_CUDA_TRY(h, cudaMemcpyAsync(h->d_theta.p, theta, (size_t)n_theta * sizeof(double), cudaMemcpyHostToDevice, s));
    int rc = enqueue_eval(h, h->d_theta.p, h->d_loss.p, grad ? h->d_grad.p : nullptr, nullptr, s);
    if (rc != UDE_OK) return rc;
    UDE_CUDA_TRY(h, cudaMemcpyAsync(loss, h->d_loss.p, (size_t)h->n_models * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (grad) UDE_CUDA_TRY(h, cudaMemcpyAsync(grad, h->d_grad.p, (size_t)n_theta * sizeof(double), cudaMemcpyDeviceToHost, s));
    UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
    h->last_h2d = 8 * n_theta; h->last_d2h = 8 * ((grad ? n_theta : 0) + h->n_models);
  }
  for (int64_t m = 0; m < h->n_models; ++m)
    if (!std::isfinite(loss[m])) return fail(h, UDE_NONFINITE, "loss is not finite (diverged trajectory?)");
  return UDE_OK;
}

const char* ude_kernel_name(ude_handle* h) {
  if (!h) return "";
  if (h->dirty && build_plan(h) != UDE_OK) return "";
  return h->k_grad ? h->k_grad->name : "";
}

int ude_forward(ude_handle* h, const double* theta, int64_t n_theta, double* out) {
  if (!h) return UDE_ERR_INVALID;
  if (!theta || !out) return fail(h, UDE_ERR_INVALID, "null argument");
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data has not been called");
  if (n_theta != h->n_theta) return fail(h, UDE_ERR_INVALID, "n_theta mismatch");
  if (h->desc.loss_kind == UDE_LOSS_DERIV_MATCH) return fail(h, UDE_ERR_INVALID, "derivative matching has no trajectories");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t n = (size_t)h->desc.d * h->n_cols;
  UDE_CUDA_TRY(h, h->d_traj.reserve(n));
  UDE_CUDA_TRY(h, cudaMemsetAsync(h->d_traj.p, 0, n * sizeof(double), s));
  UDE_CUDA_TRY(h, cudaMemcpyAsync(h->d_theta.p, theta, (size_t)n_theta * sizeof(double), cudaMemcpyHostToDevice, s));
  int rc = enqueue_eval(h, h->d_theta.p, h->d_loss.p, nullptr, h->d_traj.p, s);
  if (rc != UDE_OK) return rc;
  UDE_CUDA_TRY(h, cudaMemcpyAsync(out, h->d_traj.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
  return UDE_OK;
}

int ude_forecast_chain(ude_handle* h, const double* theta, int64_t n_theta, const double* u0, const double* limits, double l,
                       double rho, double* out) {
  if (!h) return UDE_ERR_INVALID;
  if (!theta || !u0 || !limits || !out) return fail(h, UDE_ERR_INVALID, "null argument");
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data has not been called");
  if (n_theta != h->n_theta) return fail(h, UDE_ERR_INVALID, "n_theta mismatch");
  if (h->n_models != 1 || h->desc.loss_kind != UDE_LOSS_STATE_SPACE) return fail(h, UDE_ERR_INVALID, "forecast chain: one model, state-space layout");
  for (int64_t s = 0; s < h->n_series; ++s)
    if (h->lengths[s] != 2) return fail(h, UDE_ERR_INVALID, "forecast chain: every series must be one interval (two columns)");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  if (h->dirty) { int rc = build_plan(h); if (rc != UDE_OK) return rc; }
  cudaStream_t s = h->stream;
  const int d = h->desc.d, SPW = h->k_fwd->spw;
  const size_t n = (size_t)d * h->n_cols;
  DevBuf<double> d_x, d_lim, d_out;
  UDE_CUDA_TRY(h, h->d_traj.reserve(n));
  UDE_CUDA_TRY(h, d_x.upload(u0, (size_t)d, s));
  UDE_CUDA_TRY(h, d_lim.upload(limits, (size_t)3 * d, s));
  UDE_CUDA_TRY(h, d_out.reserve((size_t)d * h->n_series));
  UDE_CUDA_TRY(h, cudaMemcpyAsync(h->d_theta.p, theta, (size_t)n_theta * sizeof(double), cudaMemcpyHostToDevice, s));
  for (int64_t j = 0; j < h->n_series; ++j) {
    ude_chain_set<<<1, 32, 0, s>>>(d_x.p, h->d_theta.p + (size_t)d * 2 * j, d);
    // the task that holds interval j (its other intervals are recomputed when their turn comes)
    int rc = launch_segments(h, h->d_theta.p, nullptr, h->d_traj.p, j / SPW, j / SPW + 1, 0, s);
    if (rc != UDE_OK) return rc;
    ude_chain_damp<<<1, 32, 0, s>>>(d_x.p, h->d_traj.p + (size_t)d * (2 * j + 1), d_lim.p, d, l, rho, d_out.p + (size_t)d * j);
  }
  UDE_CUDA_TRY(h, cudaGetLastError());
  UDE_CUDA_TRY(h, cudaMemcpyAsync(out, d_out.p, (size_t)d * h->n_series * sizeof(double), cudaMemcpyDeviceToHost, s));
  UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
  return UDE_OK;
}

int ude_comm_unique_id(void* id_bytes, int32_t n_bytes) {
  std::string err;
  if (!id_bytes || n_bytes < 128) return fail(nullptr, UDE_ERR_INVALID, "id buffer must hold 128 bytes");
  if (!nccl_load(err)) return fail(nullptr, UDE_ERR_NCCL, err);
  Id128 id; memset(&id, 0, sizeof id);
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail(nullptr, UDE_ERR_NCCL, "ncclGetUniqueId failed");
  memcpy(id_bytes, &id, 128);
  return UDE_OK;
}

int ude_comm_init(ude_handle* h, const void* id_bytes, int32_t n_bytes, int32_t n_ranks, int32_t rank) {
  if (!h) return UDE_ERR_INVALID;
  if (!id_bytes || n_bytes < 128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(h, UDE_ERR_INVALID, "bad communicator arguments");
  if (n_ranks == 1) { h->n_ranks = 1; h->rank = 0; return UDE_OK; }
  std::string err;
  if (!nccl_load(err)) return fail(h, UDE_ERR_NCCL, err);
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  Id128 id; memcpy(&id, id_bytes, 128);
  int rc = g_nccl.CommInitRank(&h->comm, n_ranks, id, rank);
  if (rc != 0) return fail(h, UDE_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  h->n_ranks = n_ranks; h->rank = rank;
  return UDE_OK;
}

struct PeerBlob {            // UDE_PEER_BLOB_BYTES
  unsigned long long magic;
  long long pid;
  int device, n_ranks;
  unsigned long long ptr, bytes;
  cudaIpcMemHandle_t ipc;    // 64 B
  char pad[UDE_PEER_BLOB_BYTES - 8 - 8 - 8 - 16 - 64];
};
static_assert(sizeof(PeerBlob) == UDE_PEER_BLOB_BYTES, "peer blob layout");
static const unsigned long long kPeerMagic = 0x5544455f50454552ull;   // "UDE_PEER"

int ude_peer_export(ude_handle* h, int32_t n_ranks, void* blob, int32_t n_bytes) {
  if (!h) return UDE_ERR_INVALID;
  if (!blob || n_bytes < UDE_PEER_BLOB_BYTES || n_ranks < 2 || n_ranks > UDE_MAX_PEERS) return fail(h, UDE_ERR_INVALID, "bad peer arguments (2 <= n_ranks <= 16, 128-byte blob)");
  if (h->n_models != 1 && h->have_data) return fail(h, UDE_ERR_INVALID, "many-model handles shard by model and need no exchange");
  if (h->x_local) return fail(h, UDE_ERR_STATE, "ude_peer_export was already called on this handle");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  const long long n_shared = h->desc.has_series_param ? h->desc.off_series_param : h->desc.n_pm;
  const int nb = (int)((n_shared + 1 + 31) / 32);      // identical on every rank (the per-series tail is not exchanged)
  h->x_ranks = n_ranks; h->x_blocks = nb; h->x_stride = (long long)nb * 32;
  h->x_bytes = 8ull * (2ull * n_ranks * (size_t)h->x_stride + 2ull * n_ranks * (size_t)nb + (size_t)nb + 1);
  UDE_CUDA_TRY(h, cudaMalloc(&h->x_local, h->x_bytes));
  UDE_CUDA_TRY(h, cudaMemset(h->x_local, 0, h->x_bytes));
  UDE_CUDA_TRY(h, cudaDeviceSynchronize());
  PeerBlob b; memset(&b, 0, sizeof b);
  b.magic = kPeerMagic; b.pid = (long long)getpid(); b.device = h->device; b.n_ranks = n_ranks;
  b.ptr = (unsigned long long)h->x_local; b.bytes = h->x_bytes;
  UDE_CUDA_TRY(h, cudaIpcGetMemHandle(&b.ipc, h->x_local));
  memcpy(blob, &b, sizeof b);
  return UDE_OK;
}

int ude_peer_attach(ude_handle* h, const void* blobs, int32_t n_ranks, int32_t rank) {
  if (!h) return UDE_ERR_INVALID;
  if (!blobs || !h->x_local || n_ranks != h->x_ranks || rank < 0 || rank >= n_ranks) return fail(h, UDE_ERR_INVALID, "ude_peer_attach: call ude_peer_export first, with the same n_ranks");
  if (h->x_attached) return fail(h, UDE_ERR_STATE, "peers are already attached");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  const PeerBlob* B = (const PeerBlob*)blobs;
  if (B[rank].magic != kPeerMagic || B[rank].ptr != (unsigned long long)h->x_local || B[rank].pid != (long long)getpid())
    return fail(h, UDE_ERR_INVALID, "ude_peer_attach: blobs[rank] is not this handle's export");
  for (int r = 0; r < n_ranks; ++r) {
    const PeerBlob& b = B[r];
    if (b.magic != kPeerMagic || b.n_ranks != n_ranks || b.bytes != h->x_bytes) return fail(h, UDE_ERR_INVALID, "ude_peer_attach: peer blob mismatch (different model or world size)");
    if (r == rank) { h->x_peer[r] = h->x_local; continue; }
    if (b.pid == (long long)getpid()) {          // a thread of this process driving another GPU
      if (b.device == h->device) return fail(h, UDE_ERR_INVALID, "two ranks on one GPU are not supported");
      int can = 0; cudaDeviceCanAccessPeer(&can, h->device, b.device);
      if (!can) return fail(h, UDE_ERR_CUDA, "no peer access between the ranks' GPUs");
      cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(h, UDE_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
      h->x_peer[r] = (void*)b.ptr;
    } else {
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { cudaGetLastError(); return fail(h, UDE_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
      h->x_peer[r] = p; h->x_opened[r] = true;
    }
  }
  h->n_ranks = n_ranks; h->rank = rank; h->x_attached = true; h->dirty = true;
  return UDE_OK;
}

int ude_ukf_loss_grad(ude_handle* h, const ude_ukf_opts* opts, const double* theta, int64_t n_theta, const double* pnu_factor,
                      double* loss, double* grad_theta, double* grad_pnu_factor) {
  if (!h) return UDE_ERR_INVALID;
  if (!opts || !theta || !pnu_factor || !loss) return fail(h, UDE_ERR_INVALID, "null argument");
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data has not been called");
  if (n_theta != h->n_theta) return fail(h, UDE_ERR_INVALID, "n_theta mismatch");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  return ukf_eval(h, opts, theta, pnu_factor, loss, grad_theta, grad_pnu_factor, nullptr, nullptr);
}

int ude_ukf_filter(ude_handle* h, const ude_ukf_opts* opts, const double* theta, int64_t n_theta, const double* pnu_factor,
                   double* states, double* covariances) {
  if (!h) return UDE_ERR_INVALID;
  if (!opts || !theta || !pnu_factor || !states) return fail(h, UDE_ERR_INVALID, "null argument");
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data has not been called");
  if (n_theta != h->n_theta) return fail(h, UDE_ERR_INVALID, "n_theta mismatch");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  double loss = 0.0;
  return ukf_eval(h, opts, theta, pnu_factor, &loss, nullptr, nullptr, states, covariances);
}

int ude_adam(ude_handle* h, double* theta, int64_t n_theta, double step_size, int32_t n_iter, double* loss_trace) {
  if (!h) return UDE_ERR_INVALID;
  if (!theta || n_iter < 0) return fail(h, UDE_ERR_INVALID, "bad argument");
  if (!h->have_data) return fail(h, UDE_ERR_STATE, "ude_set_data has not been called");
  if (n_theta != h->n_theta) return fail(h, UDE_ERR_INVALID, "n_theta mismatch");
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t n = (size_t)n_theta, M = (size_t)h->n_models;
  UDE_CUDA_TRY(h, h->d_adam_m.reserve(n));
  UDE_CUDA_TRY(h, h->d_adam_v.reserve(n));
  UDE_CUDA_TRY(h, cudaMemsetAsync(h->d_adam_m.p, 0, n * sizeof(double), s));
  UDE_CUDA_TRY(h, cudaMemsetAsync(h->d_adam_v.p, 0, n * sizeof(double), s));
  DevBuf<double> d_trace;
  UDE_CUDA_TRY(h, d_trace.reserve(M * (size_t)std::max(n_iter, 1)));
  UDE_CUDA_TRY(h, cudaMemcpyAsync(h->d_theta.p, theta, n * sizeof(double), cudaMemcpyHostToDevice, s));
  double b1t = 1.0, b2t = 1.0;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 8);
  if (h->dirty) { int rc = build_plan(h); if (rc != UDE_OK) return rc; }
  // ---- small models are launch bound (4-5 launches of a few microseconds per iteration): capture ONE iteration in a
  // CUDA graph and replay it.  Falls through to the plain loop with a communicator, when profiling, or if capture fails.
  const char* genv = getenv("UDE_ADAM_GRAPH");
  bool graphed = false;
  if (n_iter > 1 && (!h->comm || h->x_attached) && h->ev_cap == 0 && !(genv && atoi(genv) == 0)) {
    DevBuf<double> d_state;
    UDE_CUDA_TRY(h, d_state.reserve(4));
    const double st0[4] = {1.0, 1.0, 0.0, 0.0};
    UDE_CUDA_TRY(h, cudaMemcpyAsync(d_state.p, st0, sizeof st0, cudaMemcpyHostToDevice, s));
    UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      ok = enqueue_eval(h, h->d_theta.p, h->d_loss.p, h->d_grad.p, nullptr, s) == UDE_OK;
      if (ok) {
        ude_adam_update_dev<<<blocks, 256, 0, s>>>(h->d_theta.p, h->d_adam_m.p, h->d_adam_v.p, h->d_grad.p, (long long)n, step_size, d_state.p);
        ude_adam_advance<<<1, 128, 0, s>>>(d_state.p, h->d_loss.p, d_trace.p, (int)M);
      }
      const bool ended = cudaStreamEndCapture(s, &graph) == cudaSuccess;
      ok = ok && ended && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    }
    if (ok) {
      for (int it = 0; it < n_iter && ok; ++it) ok = cudaGraphLaunch(exec, s) == cudaSuccess;
      graphed = ok;
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (!graphed) {            // leave no sticky capture error behind and restart from the caller's theta
      cudaGetLastError();
      UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
      UDE_CUDA_TRY(h, cudaMemsetAsync(h->d_adam_m.p, 0, n * sizeof(double), s));
      UDE_CUDA_TRY(h, cudaMemsetAsync(h->d_adam_v.p, 0, n * sizeof(double), s));
      UDE_CUDA_TRY(h, cudaMemcpyAsync(h->d_theta.p, theta, n * sizeof(double), cudaMemcpyHostToDevice, s));
    } else {
      UDE_CUDA_TRY(h, cudaStreamSynchronize(s));      // d_state must outlive the replays
    }
  }
  h->adam_used_graph = graphed ? 1 : 0;
  for (int it = 0; it < n_iter && !graphed; ++it) {
    int rc = enqueue_eval(h, h->d_theta.p, d_trace.p + (size_t)it * M, h->d_grad.p, nullptr, s);
    if (rc != UDE_OK) return rc;
    b1t *= 0.9; b2t *= 0.999;
    ude_adam_update<<<blocks, 256, 0, s>>>(h->d_theta.p, h->d_adam_m.p, h->d_adam_v.p, h->d_grad.p, (long long)n, step_size, 1.0 - b1t, 1.0 - b2t);
    UDE_CUDA_TRY(h, cudaGetLastError());
  }
  UDE_CUDA_TRY(h, cudaMemcpyAsync(theta, h->d_theta.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (loss_trace && n_iter > 0) UDE_CUDA_TRY(h, cudaMemcpyAsync(loss_trace, d_trace.p, M * (size_t)n_iter * sizeof(double), cudaMemcpyDeviceToHost, s));
  std::vector<double> last(M, 0.0);
  if (n_iter > 0) UDE_CUDA_TRY(h, cudaMemcpyAsync(last.data(), d_trace.p + (size_t)(n_iter - 1) * M, M * sizeof(double), cudaMemcpyDeviceToHost, s));
  UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
  // a diverged fit, or a peer that never arrived at the exchange (the fused exchange then writes NaN), must not look like success
  for (size_t m = 0; m < M && n_iter > 0; ++m)
    if (!std::isfinite(last[m])) return fail(h, UDE_NONFINITE, "ude_adam: the loss of the last iteration is not finite (diverged fit, or a peer timed out in the exchange)");
  return UDE_OK;
}

int ude_adam_used_graph(const ude_handle* h) { return h ? h->adam_used_graph : 0; }

int ude_profile_begin(ude_handle* h, int32_t max_evals) {
  if (!h || max_evals < 0) return UDE_ERR_INVALID;
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  while ((int)h->ev0.size() < max_evals) {
    cudaEvent_t a, b;
    UDE_CUDA_TRY(h, cudaEventCreate(&a)); UDE_CUDA_TRY(h, cudaEventCreate(&b));
    h->ev0.push_back(a); h->ev1.push_back(b);
  }
  h->ev_used = 0; h->ev_cap = max_evals;
  return UDE_OK;
}

int ude_profile_read(ude_handle* h, float* ms_out, int32_t capacity) {
  if (!h || !ms_out) return UDE_ERR_INVALID;
  UDE_CUDA_TRY(h, cudaSetDevice(h->device));
  const int n = std::min(h->ev_used, (int)capacity);
  for (int i = 0; i < n; ++i) {
    UDE_CUDA_TRY(h, cudaEventSynchronize(h->ev1[i]));
    UDE_CUDA_TRY(h, cudaEventElapsedTime(&ms_out[i], h->ev0[i], h->ev1[i]));
  }
  h->ev_cap = 0;
  return n;
}

static int measure_fma_peak(int32_t device, double* tflops, bool f32) {
  if (!tflops) return UDE_ERR_INVALID;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return fail(nullptr, UDE_ERR_NO_DEVICE, "no such CUDA device"); }
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, UDE_ERR_CUDA, cudaGetErrorString(e));
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device);
  double* d_out = nullptr; cudaMalloc((void**)&d_out, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  float best = 1e30f;
  for (int r = 0; r < 6; ++r) {
    cudaEventRecord(e0);
    if (f32) ude_fma_peak<float><<<blocks, threads>>>((float*)d_out, iters, 1.0000001f, 1e-9f);
    else ude_fma_peak<double><<<blocks, threads>>>(d_out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
  if (e != cudaSuccess) return fail(nullptr, UDE_ERR_CUDA, cudaGetErrorString(e));
  *tflops = 2.0 * 8 * (double)iters * blocks * threads / (best * 1e-3) * 1e-12;
  return UDE_OK;
}

int ude_measure_fp64_peak(int32_t device, double* tflops) { return measure_fma_peak(device, tflops, false); }
int ude_measure_fp32_peak(int32_t device, double* tflops) { return measure_fma_peak(device, tflops, true); }

}  // extern "C"
#pragma GCC visibility pop
