// How well do DMMA phases and DFMA-chain (tanh) phases of different warps share the FP64 pipe of one SM sub-partition?
// 3 blocks of 128 threads per SM (the occupancy of ude_mma2_kernel): every scheduler holds three warps, one per block.
//   mode 0: every warp runs DMMA only            mode 1: every warp runs tanh only
//   mode 2: every warp alternates 16 DMMA / 8 tanh (the forward stage of the C3 kernel)
//   mode 3: blocks 0,1 (mod 3) run DMMA only, block 2 runs tanh only (fixed roles: two MMA warps + one tanh warp per scheduler)
//   mode 4: like 2 but the tanh levels are interleaved between the DMMAs of the same warp (fine-grained mix)
// Output: for each mode the FP64-pipe time the issued work needs at the measured peaks (16 cycles per DMMA, 2.18 per
// FP64 instruction) divided by the elapsed cycles = achievable pipe utilisation for that mix.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../universaldiffeq.jl_b200/csrc/ude_math.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int MODE, int TB>
__global__ void __launch_bounds__(128, 3) k_mix(double* out, int iters) {
  __shared__ double s_tab[64];
  ude_load_exp2_tab(s_tab);
  __syncthreads();
  double acc[4][2], z[8];
  const double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
#pragma unroll
  for (int c = 0; c < 4; ++c) { acc[c][0] = c; acc[c][1] = -c; }
#pragma unroll
  for (int c = 0; c < 8; ++c) z[c] = (threadIdx.x % 37) * 0.05 - 0.9 + 0.01 * c;
  const int role = (MODE == 3) ? (blockIdx.x % 3 == 2 ? 1 : 0) : MODE;     // 0 DMMA only, 1 tanh only, 2 both
  for (int i = 0; i < iters; ++i) {
    if (role == 0 || role == 2 || role == 4) {
      if (role != 4) {
#pragma unroll
        for (int q = 0; q < 16; ++q) dmma(acc[q & 3], a, b);
      }
    }
    if (role == 1 || role == 2) {
#pragma unroll
      for (int bb = 0; bb < 8; bb += TB) {
        double zi[TB], ho[TB];
#pragma unroll
        for (int q = 0; q < TB; ++q) zi[q] = z[bb + q] * 1.7;
        ude_tanh_fast_vec<TB>(zi, ho, s_tab);
#pragma unroll
        for (int q = 0; q < TB; ++q) z[bb + q] = ho[q] + 0.3;
      }
    }
    if (role == 4) {
      // 8 tanh as two batches of 4; 8 DMMAs issued before each batch and (by ptxas scheduling across the volatile asm
      // boundaries only partly) overlapped with it
#pragma unroll
      for (int bb = 0; bb < 8; bb += 4) {
        double zi[4], ho[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) zi[q] = z[bb + q] * 1.7;
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma(acc[q & 3], a, b);
        ude_tanh_fast_vec<4>(zi, ho, s_tab);
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma(acc[q & 3], a, b);
#pragma unroll
        for (int q = 0; q < 4; ++q) z[bb + q] = ho[q] + 0.3;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) s += acc[c][0] + acc[c][1];
#pragma unroll
  for (int c = 0; c < 8; ++c) s += z[c];
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
  const int sms = p.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  double* d_out; CK(cudaMalloc(&d_out, 64));
  const int iters = 1 << 11;
  printf("{\"gpu\": \"%s\", \"blocks_per_sm\": 3, \"warps_per_scheduler\": 3", p.name);
  // per warp-iteration: 16 DMMA = 256 pipe cycles; 8 tanh = 8 * 16 FP64 instructions (14 + scale + offset) * 2.18 = 279
  auto report = [&](const char* name, float ms, double dmma_per_it, double fp64_per_it, double warps_per_sched) {
    const double cyc = ms * 1e-3 * clk_khz * 1e3;
    const double need = iters * warps_per_sched * (dmma_per_it * 16.0 + fp64_per_it * 2.18);
    printf(", \"%s_ms\": %.4f, \"%s_pipe_util\": %.3f", name, ms, name, need / cyc);
  };
  report("dmma_only", time_ms([&] { k_mix<0, 4><<<sms * 3, 128>>>(d_out, iters); }), 16, 0, 3);
  report("tanh_only_tb4", time_ms([&] { k_mix<1, 4><<<sms * 3, 128>>>(d_out, iters); }), 0, 128, 3);
  report("tanh_only_tb8", time_ms([&] { k_mix<1, 8><<<sms * 3, 128>>>(d_out, iters); }), 0, 128, 3);
  report("tanh_only_tb2", time_ms([&] { k_mix<1, 2><<<sms * 3, 128>>>(d_out, iters); }), 0, 128, 3);
  report("alternate_tb4", time_ms([&] { k_mix<2, 4><<<sms * 3, 128>>>(d_out, iters); }), 16, 128, 3);
  report("alternate_tb8", time_ms([&] { k_mix<2, 8><<<sms * 3, 128>>>(d_out, iters); }), 16, 128, 3);
  {
    float ms = time_ms([&] { k_mix<3, 4><<<sms * 3, 128>>>(d_out, iters); });
    const double cyc = ms * 1e-3 * clk_khz * 1e3;
    const double need = iters * (2 * 16 * 16.0 + 1 * 128 * 2.18);
    printf(", \"roles_2mma_1tanh_ms\": %.4f, \"roles_2mma_1tanh_pipe_util\": %.3f", ms, need / cyc);
  }
  report("interleaved", time_ms([&] { k_mix<4, 4><<<sms * 3, 128>>>(d_out, iters); }), 16, 128, 3);
  printf("}\n");
  return 0;
}
