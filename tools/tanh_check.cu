// Accuracy and throughput of the FP64 tanh variants of csrc/ude_math.cuh against libdevice tanh (<= 1-2 ulp).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tanh_check tanh_check.cu ; prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include "../universaldiffeq.jl_b200/csrc/ude_math.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void k_acc(double* maxerr, long long n, double lo, double hi) {
  __shared__ double s_tab[256];
  if (MODE == 4) ude_load_exp2_tabb256(s_tab); else if (MODE == 3) ude_load_exp2_tabb_n<1>(s_tab); else ude_load_exp2_tab(s_tab);
  __syncthreads();
  double worst = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double x[1] = {lo + (hi - lo) * ((double)i + 0.37) / (double)n}, y[1];
    if (MODE == 1) ude_tanh_fast_vec<1>(x, y, s_tab); else if (MODE == 2) ude_tanh_fast2_vec<1>(x, y, s_tab); else if (MODE == 3) ude_tanh_fast3_vec<1>(x, y, s_tab); else ude_tanh_fast4_vec<1>(x, y, s_tab);
    worst = fmax(worst, fabs(y[0] - tanh(x[0])));
  }
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)maxerr, (unsigned long long)__double_as_longlong(worst));   // positive doubles order as integers
}

template <int MODE, int N>
__global__ void k_tput(double* out, int iters, double scale) {
  __shared__ double s_tab[64 * 16];
  if (MODE == 3) ude_load_exp2_tabb_n<16>(s_tab); else ude_load_exp2_tab_n<16>(s_tab);
  __syncthreads();
  double x[N], y[N], acc = 0.0;
#pragma unroll
  for (int c = 0; c < N; ++c) x[c] = scale * (0.01 * threadIdx.x + 0.1 * c - 1.0);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1) ude_tanh_fast_vec<N, 16>(x, y, s_tab, threadIdx.x & 15);
    else if (MODE == 2) ude_tanh_fast2_vec<N, 16>(x, y, s_tab, threadIdx.x & 15);
    else ude_tanh_fast3_vec<N, 16>(x, y, s_tab, threadIdx.x & 15);
#pragma unroll
    for (int c = 0; c < N; ++c) { acc += y[c]; x[c] = x[c] * 0.999 + 1e-3 * y[c]; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F> float time_ms(F f) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms;
}

int main() {
  double* d_err; CK(cudaMalloc(&d_err, 8));
  double* d_out; CK(cudaMalloc(&d_out, 148 * 12 * 128 * 8));
  const double ranges[6][2] = {{-22.0, 22.0}, {-2.0, 2.0}, {-0.05, 0.05}, {-1e-4, 1e-4}, {-1e6, 1e6}, {-1e200, 1e200}};
  printf("{");
  for (int m = 1; m <= 4; ++m)
    for (int r = 0; r < 6; ++r) {
      CK(cudaMemset(d_err, 0, 8));
      const long long n = 1LL << 28;
      if (m == 1) k_acc<1><<<148 * 8, 256>>>(d_err, n, ranges[r][0], ranges[r][1]);
      else if (m == 2) k_acc<2><<<148 * 8, 256>>>(d_err, n, ranges[r][0], ranges[r][1]);
      else if (m == 3) k_acc<3><<<148 * 8, 256>>>(d_err, n, ranges[r][0], ranges[r][1]);
      else k_acc<4><<<148 * 8, 256>>>(d_err, n, ranges[r][0], ranges[r][1]);
      CK(cudaDeviceSynchronize());
      double e; CK(cudaMemcpy(&e, d_err, 8, cudaMemcpyDeviceToHost));
      printf("\"v%d_max_abs_err_[%g,%g]\": %.3e, ", m, ranges[r][0], ranges[r][1], e);
    }
  const int it = 2000;
  const double evals = 148.0 * 12 * 128 * it * 8;
  float t1 = time_ms([&] { k_tput<1, 8><<<148 * 12, 128>>>(d_out, it, 1.3); });
  float t2 = time_ms([&] { k_tput<2, 8><<<148 * 12, 128>>>(d_out, it, 1.3); });
  float t3 = time_ms([&] { k_tput<3, 8><<<148 * 12, 128>>>(d_out, it, 1.3); });
  printf("\"v1_gtanh_per_s\": %.2f, \"v2_gtanh_per_s\": %.2f, \"v3_gtanh_per_s\": %.2f}\n", evals / t1 * 1e-6, evals / t2 * 1e-6, evals / t3 * 1e-6);
  return 0;
}
