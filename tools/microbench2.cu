// Round-2 microbenchmarks (B200, sm_100a) behind the ude_kernel_mma2.cuh design choices:
//   * ude_tanh_fast_vec (14 FP64-pipe instructions) vs ude_tanh_tab_vec (19): throughput and max error vs libdevice tanh
//   * DMMA m8n8k4 dependent-issue latency, and DMMA throughput at the kernels' occupancies (8 / 12 warps per SM)
//   * DMMA fed by one LDS.64 per MMA vs one LDS.64 per two MMAs (shared B fragment for two M tiles)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench2 microbench2.cu ; prints one JSON object
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include "../universaldiffeq.jl_b200/csrc/ude_math.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void k_tanh_vec(double* out, int iters, double scale) {
  __shared__ double s_tab[64];
  ude_load_exp2_tab(s_tab);
  __syncthreads();
  double acc[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c] = (threadIdx.x % 37) * 0.05 - 0.9 + 0.01 * c;
  for (int i = 0; i < iters; ++i) {
    double x[4], y[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = acc[c] * scale;
    if (MODE == 0) ude_tanh_tab_vec<4>(x, y, s_tab); else ude_tanh_fast_vec<4>(x, y, s_tab);
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = y[c] + 0.3;
  }
  double s = acc[0] + acc[1] + acc[2] + acc[3];
  if (s == 123.456) out[0] = s;
}

template <int MODE>
__global__ void k_tanh_acc(double* maxerr, int n, double lo, double hi) {
  __shared__ double s_tab[64];
  ude_load_exp2_tab(s_tab);
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[2] = {lo + (hi - lo) * (double)i / (double)(n - 1), 0.0}, y[2];
  x[1] = -x[0] * 0.5;
  if (MODE == 0) ude_tanh_tab_vec<2>(x, y, s_tab); else ude_tanh_fast_vec<2>(x, y, s_tab);
  for (int k = 0; k < 2; ++k) {
    double b = tanh(x[k]);
    double e = fabs(y[k] - b);
    double rel = (fabs(x[k]) > 1e-3) ? e / fabs(b) : 0.0;
    atomicMax((unsigned long long*)&maxerr[0], (unsigned long long)__double_as_longlong(e));
    atomicMax((unsigned long long*)&maxerr[1], (unsigned long long)__double_as_longlong(rel));
  }
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma_lat(double* out, long long* cyc, int iters) {
  double c0 = 0.5, c1 = -0.5, a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b);
    dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (c0 + c1 == 123.456) out[0] = c0;
}

template <int CHAINS>
__global__ void k_dmma(double* out, int iters) {
  double c0[CHAINS], c1[CHAINS];
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { c0[c] = c; c1[c] = -c; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) dmma(c0[c], c1[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += c0[c] + c1[c];
  if (s == 123.456) out[0] = s;
}

// B fragment from shared memory: PER = MMAs fed by one LDS.64 (1: ude_mma_kernel, 2: ude_mma2_kernel with NT = 2)
template <int PER>
__global__ void k_dmma_lds(double* out, int iters) {
  __shared__ double s_b[16 * 32];
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) s_b[i] = 1.0 - i * 1e-4;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double c0[8], c1[8];
  double a = threadIdx.x * 1e-3;
#pragma unroll
  for (int c = 0; c < 8; ++c) { c0[c] = c; c1[c] = -c; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; c += PER) {
      const double b = s_b[((c + i) & 15) * 32 + lane];
#pragma unroll
      for (int q = 0; q < PER; ++q) dmma(c0[c + q], c1[c + q], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += c0[c] + c1[c];
  if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* d_out; long long* d_cyc; double* d_err;
  CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_cyc, 64)); CK(cudaMalloc(&d_err, 64));
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  const int blocks = sms * 8, threads = 256;
  {
    const int it = 1 << 10;
    float m0 = time_ms([&] { k_tanh_vec<0><<<blocks, threads>>>(d_out, it, 1.7); });
    float m1 = time_ms([&] { k_tanh_vec<1><<<blocks, threads>>>(d_out, it, 1.7); });
    double n = 4.0 * it * blocks * threads;
    printf(", \"tanh_tab_vec4_gops\": %.2f, \"tanh_fast_vec4_gops\": %.2f", n / m0 * 1e-6, n / m1 * 1e-6);
  }
  for (int mode = 0; mode < 2; ++mode) {
    CK(cudaMemset(d_err, 0, 64));
    int n = 1 << 22;
    auto run = [&](double lo, double hi) {
      if (mode == 0) k_tanh_acc<0><<<(n + 255) / 256, 256>>>(d_err, n, lo, hi); else k_tanh_acc<1><<<(n + 255) / 256, 256>>>(d_err, n, lo, hi);
    };
    run(-22.0, 22.0); run(-3.0, 3.0); run(-1.0, 1.0); run(-1e-2, 1e-2); run(-1e-5, 1e-5);
    CK(cudaDeviceSynchronize());
    double e[2]; CK(cudaMemcpy(e, d_err, 16, cudaMemcpyDeviceToHost));
    printf(", \"tanh_%s_max_abs_err\": %.3e, \"tanh_%s_max_rel_err_absx_gt_1e-3\": %.3e", mode ? "fast" : "tab", e[0], mode ? "fast" : "tab", e[1]);
  }
  {
    k_dmma_lat<<<1, 32>>>(d_out, d_cyc, 2048); CK(cudaDeviceSynchronize());
    k_dmma_lat<<<1, 32>>>(d_out, d_cyc, 2048); CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"dmma_dependent_latency_cyc\": %.2f", (double)cyc / (2048.0 * 8));
  }
  const int iters = 1 << 12;
  auto tf = [&](float ms, int chains, int nblocks, int nthreads) { return 2.0 * 256 * chains * (double)iters * nblocks * (nthreads / 32) / ms * 1e-9; };
  {
    float a = time_ms([&] { k_dmma<8><<<blocks, threads>>>(d_out, iters); });
    printf(", \"dmma_tflops_64warps_ilp8\": %.2f", tf(a, 8, blocks, threads));
    float b = time_ms([&] { k_dmma<4><<<sms * 3, 128>>>(d_out, iters); });
    printf(", \"dmma_tflops_12warps_ilp4\": %.2f", tf(b, 4, sms * 3, 128));
    float c = time_ms([&] { k_dmma<8><<<sms * 2, 128>>>(d_out, iters); });
    printf(", \"dmma_tflops_8warps_ilp8\": %.2f", tf(c, 8, sms * 2, 128));
    float d = time_ms([&] { k_dmma<2><<<sms * 3, 128>>>(d_out, iters); });
    printf(", \"dmma_tflops_12warps_ilp2\": %.2f", tf(d, 2, sms * 3, 128));
    float e = time_ms([&] { k_dmma<4><<<sms * 2, 128>>>(d_out, iters); });
    printf(", \"dmma_tflops_8warps_ilp4\": %.2f", tf(e, 4, sms * 2, 128));
    float f = time_ms([&] { k_dmma<1><<<sms * 2, 128>>>(d_out, iters); });
    printf(", \"dmma_tflops_8warps_ilp1\": %.2f", tf(f, 1, sms * 2, 128));
  }
  {
    float a = time_ms([&] { k_dmma_lds<1><<<sms * 3, 128>>>(d_out, iters); });
    float b = time_ms([&] { k_dmma_lds<2><<<sms * 3, 128>>>(d_out, iters); });
    float c = time_ms([&] { k_dmma_lds<1><<<sms * 2, 128>>>(d_out, iters); });
    float d = time_ms([&] { k_dmma_lds<2><<<sms * 2, 128>>>(d_out, iters); });
    printf(", \"dmma_lds1_tflops_12warps\": %.2f, \"dmma_lds_per2_tflops_12warps\": %.2f, \"dmma_lds1_tflops_8warps\": %.2f, \"dmma_lds_per2_tflops_8warps\": %.2f",
           tf(a, 8, sms * 3, 128), tf(b, 8, sms * 3, 128), tf(c, 8, sms * 2, 128), tf(d, 8, sms * 2, 128));
  }
  printf("}\n");
  return 0;
}
