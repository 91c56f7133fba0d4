// B200 (sm_100a) pipe microbenchmarks that size the UDE loss+adjoint kernels.
//
// The hot path is FP64-issue bound, and MEASURED_PEAKS.json has no
// FP64 entry, so the roofline denominator has to be measured here:
//   * DFMA throughput (the "FP64 peak" every roofline.frac in bench.py divides by)
//   * DFMA dependent-issue latency (how many independent chains a warp needs)
//   * SHFL throughput alone and interleaved with DFMA (warp-shuffle matvec reductions)
//   * tanh candidates (libdevice tanh vs the expm1-based ude_tanh used by the kernels)
//   * FP64 reciprocal / division cost, DMMA m8n8k4 throughput (is the tensor path worth it?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu
// Run  : ./microbench            (prints one JSON object)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include "../universaldiffeq.jl_b200/csrc/ude_math.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- DFMA throughput
template <int CHAINS>
__global__ void k_dfma_tput(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------- DFMA latency (1 warp, 1 chain)
__global__ void k_dfma_lat(double* out, long long* cyc, int iters, double a, double b) {
  double acc = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b);
    acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b); acc = fma(acc, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
  if (acc == 123.456) out[0] = acc;
}

// ---------------------------------------------------------------- SHFL throughput
template <int CHAINS>
__global__ void k_shfl_tput(int* out, int iters) {
  int v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = __shfl_xor_sync(0xffffffffu, v[c], 1 + (c & 15));
  }
  int s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += v[c];
  if (s == 0x7fffffff) out[0] = s;
}

// DFMA + SHFL interleaved: per iteration CHAINS dfma and NSH shuffles (of a double = 2 SHFL each)
template <int CHAINS, int NSH>
__global__ void k_mix(double* out, int iters, double a, double b) {
  double acc[CHAINS];
  double sh[NSH > 0 ? NSH : 1];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
#pragma unroll
  for (int c = 0; c < (NSH > 0 ? NSH : 1); ++c) sh[c] = threadIdx.x + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
#pragma unroll
    for (int c = 0; c < NSH; ++c) sh[c] = __shfl_xor_sync(0xffffffffu, sh[c], 1 + (c & 15));
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
#pragma unroll
  for (int c = 0; c < (NSH > 0 ? NSH : 1); ++c) s += sh[c];
  if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------- tanh candidates
template <int MODE, int CHAINS>
__global__ void k_tanh(double* out, int iters, double scale) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = (threadIdx.x % 37) * 0.05 - 0.9 + 0.01 * c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      double x = acc[c] * scale;
      double y;
      if (MODE == 0) y = tanh(x);
      else if (MODE == 1) y = ude_tanh(x);
      else if (MODE == 2) y = 1.0 / (x + 2.5);
      else y = ude_rcp(x + 2.5);
      acc[c] = y + 0.3;
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  if (s == 123.456) out[0] = s;
}

// accuracy of ude_tanh vs libdevice tanh (which is <= 1-2 ulp) over a grid
__global__ void k_tanh_acc(double* maxerr, int n, double lo, double hi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = lo + (hi - lo) * (double)i / (double)(n - 1);
  double a = ude_tanh(x), b = tanh(x);
  double e = fabs(a - b);
  double rel = (b != 0.0) ? e / fabs(b) : e;
  // absolute and relative error maxima via atomicMax on the bit pattern (non-negative doubles order as ints)
  atomicMax((unsigned long long*)&maxerr[0], (unsigned long long)__double_as_longlong(e));
  atomicMax((unsigned long long*)&maxerr[1], (unsigned long long)__double_as_longlong(rel));
}

// ---------------------------------------------------------------- DMMA m8n8k4 throughput
template <int CHAINS>
__global__ void k_dmma(double* out, int iters) {
  double c0[CHAINS], c1[CHAINS];
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { c0[c] = c; c1[c] = -c; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += c0[c] + c1[c];
  if (s == 123.456) out[0] = s;
}

// DMMA and DFMA interleaved: do the FP64 tensor sub-pipe and the FP64 FMA pipe overlap or share one datapath?
template <int NMMA, int NFMA>
__global__ void k_mix_dmma_dfma(double* out, int iters, double fa, double fb) {
  double c0[NMMA > 0 ? NMMA : 1], c1[NMMA > 0 ? NMMA : 1], acc[NFMA > 0 ? NFMA : 1];
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
#pragma unroll
  for (int c = 0; c < (NMMA > 0 ? NMMA : 1); ++c) { c0[c] = c; c1[c] = -c; }
#pragma unroll
  for (int c = 0; c < (NFMA > 0 ? NFMA : 1); ++c) acc[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < NMMA; ++c)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
#pragma unroll
    for (int c = 0; c < NFMA; ++c) acc[c] = fma(acc[c], fa, fb);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < (NMMA > 0 ? NMMA : 1); ++c) s += c0[c] + c1[c];
#pragma unroll
  for (int c = 0; c < (NFMA > 0 ? NFMA : 1); ++c) s += acc[c];
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
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  double* d_out; long long* d_cyc; int* d_iout; double* d_err;
  CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_cyc, 64)); CK(cudaMalloc(&d_iout, 64)); CK(cudaMalloc(&d_err, 64));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, clk_khz);

  const int iters = 1 << 14;
  const int blocks = sms * 8, threads = 256;   // 2048 threads/SM = 64 warps/SM
  {
    float ms = time_ms([&] { k_dfma_tput<8><<<blocks, threads>>>(d_out, iters, 1.0000001, 1e-9); });
    double flops = 2.0 * 8 * (double)iters * blocks * threads;
    printf(", \"dfma_tflops\": %.3f", flops / ms * 1e-9);
    // lanes per clock per SM at nominal max clock
    printf(", \"dfma_lanes_per_clk_per_sm_at_maxclk\": %.2f", 8.0 * iters * blocks * threads / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    float ms = time_ms([&] { k_dfma_tput<1><<<blocks, threads>>>(d_out, iters, 1.0000001, 1e-9); });
    double flops = 2.0 * 1 * (double)iters * blocks * threads;
    printf(", \"dfma_tflops_ilp1_64warps\": %.3f", flops / ms * 1e-9);
  }
  for (int wps = 4; wps <= 32; wps *= 2) {     // ILP1 with few warps per SM: latency-bound regime
    float ms = time_ms([&] { k_dfma_tput<1><<<sms, wps * 32>>>(d_out, iters, 1.0000001, 1e-9); });
    double flops = 2.0 * (double)iters * sms * wps * 32;
    printf(", \"dfma_tflops_ilp1_%dwarps\": %.3f", wps, flops / ms * 1e-9);
  }
  for (int wps = 4; wps <= 16; wps *= 2) {
    float ms = time_ms([&] { k_dfma_tput<4><<<sms, wps * 32>>>(d_out, iters, 1.0000001, 1e-9); });
    double flops = 2.0 * 4 * (double)iters * sms * wps * 32;
    printf(", \"dfma_tflops_ilp4_%dwarps\": %.3f", wps, flops / ms * 1e-9);
  }
  {
    k_dfma_lat<<<1, 32>>>(d_out, d_cyc, 4096, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    k_dfma_lat<<<1, 32>>>(d_out, d_cyc, 4096, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"dfma_latency_cyc\": %.2f", (double)cyc / (4096.0 * 8));
  }
  {
    float ms = time_ms([&] { k_shfl_tput<8><<<blocks, threads>>>(d_iout, iters); });
    double n = 8.0 * iters * blocks * (threads / 32);   // warp-level SHFL instructions
    printf(", \"shfl_warpinst_per_clk_per_sm_at_maxclk\": %.3f", n / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    float ms0 = time_ms([&] { k_mix<8, 0><<<blocks, threads>>>(d_out, iters / 4, 1.0000001, 1e-9); });
    float ms2 = time_ms([&] { k_mix<8, 2><<<blocks, threads>>>(d_out, iters / 4, 1.0000001, 1e-9); });
    float ms4 = time_ms([&] { k_mix<8, 4><<<blocks, threads>>>(d_out, iters / 4, 1.0000001, 1e-9); });
    float ms8 = time_ms([&] { k_mix<8, 8><<<blocks, threads>>>(d_out, iters / 4, 1.0000001, 1e-9); });
    printf(", \"mix_8dfma_ms\": %.4f, \"mix_8dfma_2dshfl_ms\": %.4f, \"mix_8dfma_4dshfl_ms\": %.4f, \"mix_8dfma_8dshfl_ms\": %.4f", ms0, ms2, ms4, ms8);
  }
  {
    const int it = 1 << 10;
    float m0 = time_ms([&] { k_tanh<0, 4><<<blocks, threads>>>(d_out, it, 1.7); });
    float m1 = time_ms([&] { k_tanh<1, 4><<<blocks, threads>>>(d_out, it, 1.7); });
    float m2 = time_ms([&] { k_tanh<2, 4><<<blocks, threads>>>(d_out, it, 1.7); });
    float m3 = time_ms([&] { k_tanh<3, 4><<<blocks, threads>>>(d_out, it, 1.7); });
    double n = 4.0 * it * blocks * threads;
    // express as "DFMA-equivalents": time per op relative to the DFMA peak measured above is printed by the caller
    printf(", \"tanh_libdevice_gops\": %.2f, \"tanh_ude_gops\": %.2f, \"div_gops\": %.2f, \"ude_rcp_gops\": %.2f",
           n / m0 * 1e-6, n / m1 * 1e-6, n / m2 * 1e-6, n / m3 * 1e-6);
  }
  {
    CK(cudaMemset(d_err, 0, 64));
    int n = 1 << 22;
    k_tanh_acc<<<(n + 255) / 256, 256>>>(d_err, n, -22.0, 22.0);
    k_tanh_acc<<<(n + 255) / 256, 256>>>(d_err, n, -1.0, 1.0);
    k_tanh_acc<<<(n + 255) / 256, 256>>>(d_err, n, -1e-3, 1e-3);
    CK(cudaDeviceSynchronize());
    double e[2]; CK(cudaMemcpy(e, d_err, 16, cudaMemcpyDeviceToHost));
    printf(", \"ude_tanh_max_abs_err_vs_libdevice\": %.3e, \"ude_tanh_max_rel_err_vs_libdevice\": %.3e", e[0], e[1]);
  }
  {
    float ms = time_ms([&] { k_dmma<8><<<blocks, threads>>>(d_out, iters / 4); });
    double flops = 2.0 * 8 * 8 * 4 * 8.0 * (iters / 4) * blocks * (threads / 32);
    printf(", \"dmma_m8n8k4_tflops\": %.3f", flops / ms * 1e-9);
  }
  {
    const int it = iters / 4;
    float m_mma = time_ms([&] { k_mix_dmma_dfma<4, 0><<<blocks, threads>>>(d_out, it, 1.0000001, 1e-9); });
    float m_fma = time_ms([&] { k_mix_dmma_dfma<0, 32><<<blocks, threads>>>(d_out, it, 1.0000001, 1e-9); });
    float m_both = time_ms([&] { k_mix_dmma_dfma<4, 32><<<blocks, threads>>>(d_out, it, 1.0000001, 1e-9); });
    // 4 DMMA (4*256 FMA) and 32 DFMA (32*32 FMA) per warp-iteration: equal FMA counts
    printf(", \"mix_4dmma_ms\": %.4f, \"mix_32dfma_ms\": %.4f, \"mix_4dmma_32dfma_ms\": %.4f", m_mma, m_fma, m_both);
  }
  printf("}\n");
  return 0;
}
