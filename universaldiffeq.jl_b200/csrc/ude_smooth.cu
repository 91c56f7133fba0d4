// Batched regularisation smoother for the derivative-matching loss (reference call sites
// src/LossFunctionConstructors.jl:161-171, :510-522: `RegularizationSmooth(u, t, d; alg = :gcv_svd)` per state and series,
// then `DataInterpolations.derivative` at the observation times).  The grid-dependent operators come from the host
// (universaldiffeq.jl_b200/smoothing.py: Vt = right singular vectors of the d-th derivative matrix D, s = its singular values,
// M = knot derivatives of the natural cubic spline as a matrix); this kernel does the per-row work, ONE WARP PER ROW:
//   z = Vt u                                  projection on the penalised directions
//   GCV(lambda) = n R / T^2,  R = sum ((1 - f_i) z_i)^2,  T = sum (1 - f_i),  f_i = 1 / (1 + lambda^2 s_i^2)
//   lambda* = argmin over [lambda_lo, lambda_hi]: 64-point scan in log(lambda) (one or two points per lane), then the root of
//             d GCV / d log(lambda) inside the best bracket by bisection (a root is located to rounding error, a flat minimum
//             is not: two implementations that MINIMISE agree to ~1e-8 only, two that solve g = 0 agree to ~1e-14)
//   x = u - V ((1 - f) z),   dx/dt = M x
// HBM traffic is the rows in and twice the rows out; the operators (a few KB) stay in L1/L2.
#include <cuda_runtime.h>

#include <cmath>
#include <string>

#include "../../include/ude_cuda.h"

namespace {

constexpr int kMaxN = 512;          // points per series the shared-memory row buffers are sized for
constexpr int kScan = 64;

// R, T and g = R' T - 2 R T' (the numerator of d GCV / d log lambda, up to a positive factor) at x = log(lambda)
__device__ __forceinline__ void gcv_terms(const double* __restrict__ z2, const double* __restrict__ s2, int r, double x,
                                          double& R, double& T, double& g) {
  const double l2 = exp(2.0 * x);
  double Rp = 0.0, Tp = 0.0;
  R = 0.0; T = 0.0;
  for (int i = 0; i < r; ++i) {
    const double a = l2 * s2[i];
    const double om = a / (1.0 + a);          // 1 - f_i
    const double f = 1.0 / (1.0 + a);
    R = fma(om * om, z2[i], R);
    T += om;
    Rp = fma(4.0 * f * om * om, z2[i], Rp);
    Tp = fma(2.0 * f, om, Tp);
  }
  g = Rp * T - 2.0 * R * Tp;
}

__global__ void __launch_bounds__(128) ude_reg_smooth_kernel(int n, int r, long long rows, const double* __restrict__ Vt,
                                                             const double* __restrict__ s, const double* __restrict__ M,
                                                             const double* __restrict__ U, double x_lo, double x_hi,
                                                             double* __restrict__ out_u, double* __restrict__ out_du,
                                                             double* __restrict__ out_lam) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double* u = sm + (size_t)warp * (3 * (size_t)n + 2 * (size_t)r);      // [u n][x n][z r][z^2 r] (+ s^2 shared per block below)
  double* x = u + n;
  double* z = x + n;
  double* z2 = z + r;
  double* s2 = sm + (size_t)wpb * (3 * (size_t)n + 2 * (size_t)r);
  for (int i = threadIdx.x; i < r; i += blockDim.x) s2[i] = s[i] * s[i];
  __syncthreads();
  for (long long row = (long long)blockIdx.x * wpb + warp; row < rows; row += (long long)gridDim.x * wpb) {
    for (int k = lane; k < n; k += 32) u[k] = U[row * n + k];
    __syncwarp();
    for (int i = lane; i < r; i += 32) {
      double a = 0.0;
      for (int k = 0; k < n; ++k) a = fma(Vt[(size_t)i * n + k], u[k], a);
      z[i] = a; z2[i] = a * a;
    }
    __syncwarp();
    // coarse scan: lane l evaluates points l and l + 32
    double best = 1e300; int bi = 0;
    for (int q = lane; q < kScan; q += 32) {
      const double xq = x_lo + (x_hi - x_lo) * (double)q / (double)(kScan - 1);
      double R, T, g;
      gcv_terms(z2, s2, r, xq, R, T, g);
      const double v = (T > 0.0) ? (double)n * R / (T * T) : 1e300;
      if (v < best) { best = v; bi = q; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    // refine: g < 0 left of a minimum, > 0 right of it
    const double h = (x_hi - x_lo) / (double)(kScan - 1);
    double a = x_lo + h * (double)max(bi - 1, 0), b = x_lo + h * (double)min(bi + 1, kScan - 1);
    double R, T, ga, gb, xs = x_lo + h * (double)bi;
    gcv_terms(z2, s2, r, a, R, T, ga);
    gcv_terms(z2, s2, r, b, R, T, gb);
    if (ga < 0.0 && gb > 0.0) {
      for (int it = 0; it < 200 && b - a > 4e-16 * fmax(fabs(a), fabs(b)); ++it) {
        const double m = 0.5 * (a + b);
        double gm;
        gcv_terms(z2, s2, r, m, R, T, gm);
        if (gm < 0.0) a = m; else b = m;
      }
      xs = 0.5 * (a + b);
    }
    const double l2 = exp(2.0 * xs);
    // x = u - sum_i (1 - f_i) z_i V_i
    for (int k = lane; k < n; k += 32) {
      double acc = 0.0;
      for (int i = 0; i < r; ++i) {
        const double aa = l2 * s2[i];
        acc = fma(aa / (1.0 + aa) * z[i], Vt[(size_t)i * n + k], acc);
      }
      x[k] = u[k] - acc;
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc = fma(M[(size_t)k * n + j], x[j], acc);
      out_u[row * n + k] = x[k];
      out_du[row * n + k] = acc;
    }
    if (lane == 0) out_lam[row] = exp(xs);
    __syncwarp();
  }
}

}  // namespace

int ude_fail_create(int code, const std::string& msg);      // ude_api.cu: message readable through ude_last_error(NULL)

#pragma GCC visibility push(default)
extern "C" int ude_reg_smooth(int32_t device, int32_t n, int32_t r, int64_t rows, const double* Vt, const double* s,
                              const double* M, const double* U, double lambda_lo, double lambda_hi, double* out_u,
                              double* out_dudt, double* out_lambda) {
  if (n < 3 || r < 1 || r > n || rows < 1 || !Vt || !s || !M || !U || !out_u || !out_dudt || !out_lambda || !(lambda_lo > 0.0) ||
      !(lambda_hi > lambda_lo))
    return ude_fail_create(UDE_ERR_INVALID, "ude_reg_smooth: bad arguments");
  if (n > kMaxN) return ude_fail_create(UDE_ERR_UNSUPPORTED, "ude_reg_smooth: more than 512 points per series");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return ude_fail_create(UDE_ERR_NO_DEVICE, "ude_reg_smooth: no CUDA device (libude_cuda has no CPU fallback)");
  }
  auto bad = [&](cudaError_t e) { return ude_fail_create(UDE_ERR_CUDA, std::string("ude_reg_smooth: ") + cudaGetErrorString(e)); };
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return bad(e);
  const size_t nb = (size_t)n * sizeof(double), rowsb = (size_t)rows * nb;
  double *dV = nullptr, *ds = nullptr, *dM = nullptr, *dU = nullptr, *dX = nullptr, *dD = nullptr, *dL = nullptr;
  auto release = [&]() { cudaFree(dV); cudaFree(ds); cudaFree(dM); cudaFree(dU); cudaFree(dX); cudaFree(dD); cudaFree(dL); };
  if ((e = cudaMalloc(&dV, (size_t)r * nb)) != cudaSuccess || (e = cudaMalloc(&ds, (size_t)r * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&dM, (size_t)n * nb)) != cudaSuccess || (e = cudaMalloc(&dU, rowsb)) != cudaSuccess ||
      (e = cudaMalloc(&dX, rowsb)) != cudaSuccess || (e = cudaMalloc(&dD, rowsb)) != cudaSuccess ||
      (e = cudaMalloc(&dL, (size_t)rows * sizeof(double))) != cudaSuccess) { release(); return bad(e); }
  cudaMemcpy(dV, Vt, (size_t)r * nb, cudaMemcpyHostToDevice);
  cudaMemcpy(ds, s, (size_t)r * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(dM, M, (size_t)n * nb, cudaMemcpyHostToDevice);
  cudaMemcpy(dU, U, rowsb, cudaMemcpyHostToDevice);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device);
  const int wpb = 4;
  const size_t smem = ((size_t)wpb * (3 * (size_t)n + 2 * (size_t)r) + (size_t)r) * sizeof(double);
  if (smem > 48 * 1024) cudaFuncSetAttribute(ude_reg_smooth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long want = (rows + wpb - 1) / wpb;
  const int grid = (int)std::min<long long>(want, (long long)prop.multiProcessorCount * 8);
  ude_reg_smooth_kernel<<<grid, wpb * 32, smem>>>(n, r, rows, dV, ds, dM, dU, log(lambda_lo), log(lambda_hi), dX, dD, dL);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out_u, dX, rowsb, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_dudt, dD, rowsb, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_lambda, dL, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost);
  release();
  return e == cudaSuccess ? UDE_OK : bad(e);
}
#pragma GCC visibility pop
