// Unscented-Kalman-filter marginal likelihood on top of the segment kernels.
//   reference: /root/reference/src/UnscentedKalmanFilter.jl:2-87 (sigma_points, ukf_update, ukf_likelihood, ukf_smoothing)
//              /root/reference/src/LossFunctionConstructors.jl:103-127 (UDE), :427-476 (MultiUDE)
// Included by ude_api.cu (needs ude_handle, build_plan, launch_segments).
//
// The filter is a recursion in time, parallel over series and over the 2d+1 sigma points of a step.  Every step k
// (interval k -> k+1 of each series that still has one) is
//     ukf_sigma_fwd   (thread per series)  U = chol((lambda+L) P); sigma points -> start columns of an AUXILIARY handle
//     segment kernel  (forward)            the 2d+1 one-interval predicts of every active series: the same fused RK kernels
//                                          as every other loss -- the auxiliary handle describes each sigma point as a
//                                          two-column mini-series, one "model" per step so a step is a contiguous task range
//     ukf_update_fwd  (thread per series)  mean, covariance, gain, posterior, log-likelihood term
// and the reverse sweep runs the steps backwards:
//     ukf_update_bwd  adjoint of the update -> lambda_j = dLoss/dy_j for every sigma point; written as the mini-series'
//                     TARGET column  y_j - lambda_j / 2: with weight B = I the state-space loss r'Br, r = target - F, has
//                     dLoss/dF = -2 r = lambda_j, so the ordinary loss+adjoint kernel returns J^T lambda_j in the gradient
//                     of the start column and (dF/dp)^T lambda_j in the step's process-model gradient: an exact VJP with
//                     an externally supplied seed, no kernel change
//     segment kernel  (loss + adjoint)
//     ukf_sigma_bwd   xbar = sum chibar_j, Ubar = chibar+ - chibar-, adjoint of the Cholesky factorisation -> Pbar.
// tests/test_ukf_oracle.py holds the same reverse sweep in NumPy, checked against complex-step derivatives.
#pragma once

namespace {

constexpr int UKF_MAXD = UDE_MAX_D;

struct UkfConst {
  int d, n_sig;                       // n_sig = 2d + 1
  double c, Wm0, Wc0, Wi, ll_const;   // c = lambda + L
  double reg;                         // weight of |W1|^2 + |W2|^2
  double Peta[UKF_MAXD * UKF_MAXD];   // column-major d x d; P_nu is per model (ude_ukf_plan::d_pnu)
};

struct UkfStep {
  int n_act;                 // active series in this step
  long long act0;            // offset into the active-series list / per-(step, series) storage
  long long col0;            // aux global column of the step's first mini-series
  int k;                     // interval k -> k+1 inside each series
  int blk0, n_blk;           // the step's auxiliary models: one per model of the handle that still has a series in step k
};

// one auxiliary model = (filter step, model of the handle): theta block [sigma-point columns of its series | process model]
struct UkfBlock {
  long long base;            // aux theta offset of the block
  long long pm_off;          // ... of its process-model part
  int model;                 // model of the handle
};

struct ude_ukf_plan {
  ude_handle* aux = nullptr;
  int d = 0, K = 0;
  std::vector<UkfStep> steps;
  int64_t M = 1, NB = 0;                // models of the handle, auxiliary models
  std::vector<UkfBlock> blocks;
  DevBuf<int> d_active, d_next;         // d_next[act0 + a]: index of the same series in the next step's list, or -1
  DevBuf<long long> d_thoff;            // per (step, active series): aux theta offset of its 2 n_sig columns
  DevBuf<int> d_smodel;                 // model of each series
  DevBuf<UkfBlock> d_blk;
  DevBuf<int> d_mblk_off, d_mblk;       // CSR: the auxiliary models of each model, in step order
  DevBuf<double> d_pnu;                 // P_nu of each model, d x d column-major
  DevBuf<double> d_xprev, d_Pprev, d_U;               // per (step, active series)
  DevBuf<double> d_x, d_P, d_xb, d_Pb, d_ll, d_pnub;  // per series
  DevBuf<double> d_theta_aux, d_grad_aux, d_traj_aux;
  DevBuf<double> d_pm, d_out, d_states, d_covs;
  DevBuf<UkfConst> d_const;
  cudaGraphExec_t exec[4] = {nullptr, nullptr, nullptr, nullptr};   // loss / loss+grad / filter / filter+covariances
  int used_graph = 0;
  int64_t n_aux_theta = 0, n_aux_cols = 0;
  ~ude_ukf_plan() {
    for (cudaGraphExec_t e : exec) if (e) cudaGraphExecDestroy(e);
    if (aux) ude_destroy(aux);
  }
};

// the plan caches host copies of times / knots, buffers sized by n_series / n_cols and captured graphs with baked-in
// pointers: it must not survive ude_set_data / ude_set_covariates / a change of the prediction window
void ukf_invalidate(ude_handle* h) {
  if (h->ukf) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    delete h->ukf;
    h->ukf = nullptr;
  }
}

// ---- small dense algebra, compile-time D <= 8 (everything stays in registers), column-major M[i + D*j] ---------------------------------------------------
template <int d>
__device__ __forceinline__ void ukf_chol_upper(const double* A, double* U) {      // reads the upper triangle of A
  #pragma unroll
  for (int i = 0; i < d * d; ++i) U[i] = 0.0;
  #pragma unroll
  for (int i = 0; i < d; ++i) {
    double s = A[i + d * i];
    #pragma unroll
    for (int k = 0; k < i; ++k) s -= U[k + d * i] * U[k + d * i];
    const double p = sqrt(s);            // NaN for a non-positive pivot: the caller sees a non-finite loss
    U[i + d * i] = p;
    #pragma unroll
    for (int j = i + 1; j < d; ++j) {
      double t = A[i + d * j];
      #pragma unroll
      for (int k = 0; k < i; ++k) t -= U[k + d * i] * U[k + d * j];
      U[i + d * j] = t / p;
    }
  }
}

// Ab (upper entries; the factorisation never reads the lower triangle) from Ub; Ub is destroyed
template <int d>
__device__ __forceinline__ void ukf_chol_upper_adjoint(const double* U, double* Ub, double* Ab) {
  #pragma unroll
  for (int i = 0; i < d * d; ++i) Ab[i] = 0.0;
  #pragma unroll
  for (int i = d - 1; i >= 0; --i) {
    const double p = U[i + d * i];
    #pragma unroll
    for (int j = d - 1; j > i; --j) {
      const double sb = Ub[i + d * j] / p;
      Ub[i + d * i] -= Ub[i + d * j] * U[i + d * j] / p;
      Ab[i + d * j] = sb;
      #pragma unroll
      for (int k = 0; k < i; ++k) { Ub[k + d * i] -= sb * U[k + d * j]; Ub[k + d * j] -= sb * U[k + d * i]; }
    }
    const double sb = Ub[i + d * i] / (2.0 * p);
    Ab[i + d * i] = sb;
    #pragma unroll
    for (int k = 0; k < i; ++k) Ub[k + d * i] -= 2.0 * sb * U[k + d * i];
  }
}

// G = inv(S) by Gauss-Jordan WITHOUT pivoting (S = P1 + P_eta is symmetric positive definite while the filter is healthy;
// a lost definiteness shows up as a non-finite loss); returns det(S) (the reference: inv(S), log(det(S))).  No pivot
// search means no data-dependent indexing, so the matrices stay in registers.
template <int d>
__device__ __forceinline__ double ukf_inverse(const double* S, double* G) {
  double M[d * d];
  #pragma unroll
  for (int i = 0; i < d * d; ++i) { M[i] = S[i]; G[i] = 0.0; }
  #pragma unroll
  for (int i = 0; i < d; ++i) G[i + d * i] = 1.0;
  double det = 1.0;
  #pragma unroll
  for (int c = 0; c < d; ++c) {
    const double p = M[c + d * c];
    det *= p;
    const double ip = 1.0 / p;
    #pragma unroll
    for (int j = 0; j < d; ++j) { M[c + d * j] *= ip; G[c + d * j] *= ip; }
    #pragma unroll
    for (int r = 0; r < d; ++r) {
      if (r == c) continue;
      const double f = M[r + d * c];
      #pragma unroll
      for (int j = 0; j < d; ++j) { M[r + d * j] -= f * M[c + d * j]; G[r + d * j] -= f * G[c + d * j]; }
    }
  }
  return det;
}

// the quantities of one update, recomputed identically by the forward and the reverse kernel
template <int d>
struct UkfUpdate {
  double m[d], r[d];
  double P1[d * d], G[d * d], Kg[d * d];
  double det;
};

// Y: n_sig predictions, Y[j] at traj[d * (col_first + 2 j + 1) + i]
template <int d>
__device__ __forceinline__ void ukf_update_common(const UkfConst& C, const double* __restrict__ pnu, const double* __restrict__ traj,
                                                  long long col_first, const double* __restrict__ y, UkfUpdate<d>& W) {
  #pragma unroll
  for (int i = 0; i < d; ++i) {
    double s = C.Wm0 * traj[d * (col_first + 1) + i];
    #pragma unroll
    for (int j = 1; j < (2 * d + 1); ++j) s += C.Wi * traj[d * (col_first + 2 * j + 1) + i];
    W.m[i] = s;
  }
  #pragma unroll
  for (int a = 0; a < d; ++a)
    #pragma unroll
    for (int b = 0; b < d; ++b) {
      double s = 0.0;
      #pragma unroll
      for (int j = 0; j < (2 * d + 1); ++j) {
        const double w = j == 0 ? C.Wc0 : C.Wi;
        const double ea = traj[d * (col_first + 2 * j + 1) + a] - W.m[a], eb = traj[d * (col_first + 2 * j + 1) + b] - W.m[b];
        s += w * ea * eb;
      }
      W.P1[a + d * b] = s + pnu[a + d * b];
    }
  double S[d * d];
  #pragma unroll
  for (int i = 0; i < d * d; ++i) S[i] = W.P1[i] + C.Peta[i];
  W.det = ukf_inverse<d>(S, W.G);
  #pragma unroll
  for (int a = 0; a < d; ++a)
    #pragma unroll
    for (int b = 0; b < d; ++b) {
      double s = 0.0;
      #pragma unroll
      for (int k = 0; k < d; ++k) s += W.P1[a + d * k] * W.G[k + d * b];
      W.Kg[a + d * b] = s;
    }
  #pragma unroll
  for (int i = 0; i < d; ++i) W.r[i] = y[i] - W.m[i];
}

__global__ void ukf_init(int n_series, int d, const long long* __restrict__ starts, const double* __restrict__ data,
                         const UkfConst* __restrict__ Cp, double* __restrict__ X, double* __restrict__ P, double* __restrict__ states,
                         double* __restrict__ covs) {
  const UkfConst& C = *Cp;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_series) return;
  const long long c0 = starts[s];
  for (int i = 0; i < d; ++i) { X[(long long)s * d + i] = data[d * c0 + i]; if (states) states[d * c0 + i] = data[d * c0 + i]; }
  for (int i = 0; i < d * d; ++i) { P[(long long)s * d * d + i] = C.Peta[i]; if (covs) covs[(long long)d * d * c0 + i] = 0.0; }
}

// every auxiliary model gets the process model of the model it belongs to
__global__ void ukf_bcast_pm(const double* __restrict__ pm, long long n_pm, const UkfBlock* __restrict__ blk, int NB,
                             double* __restrict__ theta_aux) {
  const int b = blockIdx.x;
  if (b >= NB) return;
  double* dst = theta_aux + blk[b].pm_off;
  const double* src = pm + (long long)blk[b].model * n_pm;
  for (long long i = threadIdx.x; i < n_pm; i += blockDim.x) dst[i] = src[i];
}

template <int d>
__device__ __forceinline__ void ukf_sigma_fwd_body(const UkfStep& st, int a, const UkfConst& C, const int* active, const long long* thoff,
                              const double* X, const double* P, double* xprev, double* Pprev,
                              double* Usave, double* theta_aux) {
  const int s = active[st.act0 + a];
  double x[d], A[d * d], U[d * d];
  #pragma unroll
  for (int i = 0; i < d; ++i) { x[i] = X[(long long)s * d + i]; xprev[(st.act0 + a) * d + i] = x[i]; }
  #pragma unroll
  for (int i = 0; i < d * d; ++i) {
    const double p = P[(long long)s * d * d + i];
    Pprev[(st.act0 + a) * d * d + i] = p;
    A[i] = C.c * p;
  }
  ukf_chol_upper<d>(A, U);
  #pragma unroll
  for (int i = 0; i < d * d; ++i) Usave[(st.act0 + a) * d * d + i] = U[i];
  double* th = theta_aux + thoff[st.act0 + a];                                 // this series' 2 n_sig columns
  #pragma unroll
  for (int i = 0; i < d; ++i) th[i] = x[i];                                    // sigma 0 -> column 0
  #pragma unroll
  for (int q = 0; q < d; ++q)
    #pragma unroll
    for (int i = 0; i < d; ++i) {
      th[d * 2 * (1 + q) + i] = x[i] + U[q + d * i];                           // x + row q of U
      th[d * 2 * (1 + d + q) + i] = x[i] - U[q + d * i];
    }
}

template <int d>
__device__ __forceinline__ void ukf_update_fwd_body(const UkfStep& st, int a, const UkfConst& C, const int* active, const int* smodel,
                               const double* pnu, const long long* starts,
                               const double* data, const double* traj, double* X,
                               double* P, double* ll, double* states, double* covs) {
  const int s = active[st.act0 + a];
  const long long col = starts[s] + st.k + 1;
  UkfUpdate<d> W;
  ukf_update_common(C, pnu + (long long)smodel[s] * d * d, traj, st.col0 + 2LL * (2 * d + 1) * a, data + d * col, W);
  double q = 0.0;
  #pragma unroll
  for (int i = 0; i < d; ++i) {
    double xi = W.m[i], gi = 0.0;
    #pragma unroll
    for (int j = 0; j < d; ++j) { xi += W.Kg[i + d * j] * W.r[j]; gi += W.G[i + d * j] * W.r[j]; }
    X[(long long)s * d + i] = xi;
    if (states) states[d * col + i] = xi;
    q += W.r[i] * gi;
  }
  #pragma unroll
  for (int i = 0; i < d; ++i)
    #pragma unroll
    for (int j = 0; j < d; ++j) {
      double v = W.P1[i + d * j];
      #pragma unroll
      for (int k = 0; k < d; ++k) v -= W.Kg[i + d * k] * W.P1[k + d * j];
      P[(long long)s * d * d + i + d * j] = v;
      if (covs) covs[(long long)d * d * col + i + d * j] = v;
    }
  ll[s] += -0.5 * q - 0.5 * log(W.det) - C.ll_const;
}

// adjoint of the update for loss = -sum ll: in (xb, Pb) = d loss / d (posterior mean, covariance); out: target columns
// y_j - lambda_j / 2 and this series' share of d loss / d P_nu
template <int d>
__device__ __forceinline__ void ukf_update_bwd_body(const UkfStep& st, int a, const UkfConst& C, const int* active, const int* smodel,
                               const double* pnu, const long long* thoff, const long long* starts,
                               const double* data, const double* traj, const double* Xb,
                               const double* Pb_in, double* pnub, double* theta_aux) {
  const int s = active[st.act0 + a];
  const long long col = starts[s] + st.k + 1, cf = st.col0 + 2LL * (2 * d + 1) * a;
  UkfUpdate<d> W;
  ukf_update_common(C, pnu + (long long)smodel[s] * d * d, traj, cf, data + d * col, W);
  double xb[d], Pb[d * d], Kb[d * d], rb[d], mb[d];
  double P1b[d * d], Gb[d * d], T[d * d];
  #pragma unroll
  for (int i = 0; i < d; ++i) xb[i] = Xb[(long long)s * d + i];
  #pragma unroll
  for (int i = 0; i < d * d; ++i) Pb[i] = Pb_in[(long long)s * d * d + i];
  // Kb = xb r' - Pb P1'
  #pragma unroll
  for (int i = 0; i < d; ++i)
    #pragma unroll
    for (int j = 0; j < d; ++j) {
      double v = xb[i] * W.r[j];
      #pragma unroll
      for (int k = 0; k < d; ++k) v -= Pb[i + d * k] * W.P1[j + d * k];
      Kb[i + d * j] = v;
    }
  // rb = K' xb + 0.5 (G + G') r ;  mb = xb - rb
  #pragma unroll
  for (int i = 0; i < d; ++i) {
    double v = 0.0;
    #pragma unroll
    for (int k = 0; k < d; ++k) v += W.Kg[k + d * i] * xb[k] + 0.5 * (W.G[i + d * k] + W.G[k + d * i]) * W.r[k];
    rb[i] = v; mb[i] = xb[i] - v;
  }
  // P1b = Pb - K' Pb + Kb G' ;  Gb = P1' Kb + 0.5 r r'
  #pragma unroll
  for (int i = 0; i < d; ++i)
    #pragma unroll
    for (int j = 0; j < d; ++j) {
      double v = Pb[i + d * j], g = 0.5 * W.r[i] * W.r[j];
      #pragma unroll
      for (int k = 0; k < d; ++k) {
        v += -W.Kg[k + d * i] * Pb[k + d * j] + Kb[i + d * k] * W.G[j + d * k];
        g += W.P1[k + d * i] * Kb[k + d * j];
      }
      P1b[i + d * j] = v; Gb[i + d * j] = g;
    }
  // Sb = -G' Gb G' + 0.5 G' ;  P1b += Sb
  #pragma unroll
  for (int i = 0; i < d; ++i)
    #pragma unroll
    for (int j = 0; j < d; ++j) {
      double v = 0.0;
      #pragma unroll
      for (int k = 0; k < d; ++k) v += W.G[k + d * i] * Gb[k + d * j];
      T[i + d * j] = v;                  // G' Gb
    }
  #pragma unroll
  for (int i = 0; i < d; ++i)
    #pragma unroll
    for (int j = 0; j < d; ++j) {
      double v = 0.5 * W.G[j + d * i];
      #pragma unroll
      for (int k = 0; k < d; ++k) v -= T[i + d * k] * W.G[j + d * k];
      P1b[i + d * j] += v;
    }
  #pragma unroll
  for (int i = 0; i < d * d; ++i) pnub[(long long)s * d * d + i] += P1b[i];
  // eb_j = w_j (P1b + P1b') e_j ; mb -= sum eb_j ; yb_j = eb_j + wm_j mb
  double sumE[d];
  #pragma unroll
  for (int i = 0; i < d; ++i) sumE[i] = 0.0;
  double* th = theta_aux + thoff[st.act0 + a];
  #pragma unroll
  for (int j = 0; j < (2 * d + 1); ++j) {
    const double w = j == 0 ? C.Wc0 : C.Wi;
    #pragma unroll
    for (int i = 0; i < d; ++i) {
      double v = 0.0;
      #pragma unroll
      for (int k = 0; k < d; ++k) v += (P1b[i + d * k] + P1b[k + d * i]) * (traj[d * (cf + 2 * j + 1) + k] - W.m[k]);
      v *= w;
      sumE[i] += v;
      th[d * (2 * j + 1) + i] = v;       // eb_j parked in the target column; finished below
    }
  }
  #pragma unroll
  for (int i = 0; i < d; ++i) mb[i] -= sumE[i];
  #pragma unroll
  for (int j = 0; j < (2 * d + 1); ++j) {
    const double wm = j == 0 ? C.Wm0 : C.Wi;
    #pragma unroll
    for (int i = 0; i < d; ++i) {
      const double lam = th[d * (2 * j + 1) + i] + wm * mb[i];
      th[d * (2 * j + 1) + i] = traj[d * (cf + 2 * j + 1) + i] - 0.5 * lam;
    }
  }
}

template <int d>
__device__ __forceinline__ void ukf_sigma_bwd_body(const UkfStep& st, int a, const UkfConst& C, const int* active, const long long* thoff,
                              const double* Usave, const double* grad_aux, double* Xb, double* Pb) {
  const int s = active[st.act0 + a];
  const double* g = grad_aux + thoff[st.act0 + a];
  double U[d * d], Ub[d * d], Ab[d * d];
  #pragma unroll
  for (int i = 0; i < d * d; ++i) U[i] = Usave[(st.act0 + a) * d * d + i];
  #pragma unroll
  for (int i = 0; i < d; ++i) {
    double v = 0.0;
    #pragma unroll
    for (int j = 0; j < (2 * d + 1); ++j) v += g[d * 2 * j + i];
    Xb[(long long)s * d + i] = v;
  }
  #pragma unroll
  for (int q = 0; q < d; ++q)
    #pragma unroll
    for (int i = 0; i < d; ++i) Ub[q + d * i] = g[d * 2 * (1 + q) + i] - g[d * 2 * (1 + d + q) + i];
  ukf_chol_upper_adjoint<d>(U, Ub, Ab);
  #pragma unroll
  for (int i = 0; i < d * d; ++i) Pb[(long long)s * d * d + i] = C.c * Ab[i];
}

template <int d>
__global__ void ukf_sigma_fwd(UkfStep st, const UkfConst* __restrict__ Cp, const int* __restrict__ active, const long long* __restrict__ thoff,
                              const double* __restrict__ X, const double* __restrict__ P, double* __restrict__ xprev,
                              double* __restrict__ Pprev, double* __restrict__ Usave, double* __restrict__ theta_aux) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= st.n_act) return;
  ukf_sigma_fwd_body<d>(st, a, *Cp, active, thoff, X, P, xprev, Pprev, Usave, theta_aux);
}

template <int d>
__global__ void ukf_update_fwd(UkfStep st, const UkfConst* __restrict__ Cp, const int* __restrict__ active, const int* __restrict__ smodel,
                               const double* __restrict__ pnu, const long long* __restrict__ starts,
                               const double* __restrict__ data, const double* __restrict__ traj, double* __restrict__ X,
                               double* __restrict__ P, double* __restrict__ ll, double* __restrict__ states, double* __restrict__ covs) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= st.n_act) return;
  ukf_update_fwd_body<d>(st, a, *Cp, active, smodel, pnu, starts, data, traj, X, P, ll, states, covs);
}

template <int d>
__global__ void ukf_update_bwd(UkfStep st, const UkfConst* __restrict__ Cp, const int* __restrict__ active, const int* __restrict__ smodel,
                               const double* __restrict__ pnu, const long long* __restrict__ thoff, const long long* __restrict__ starts,
                               const double* __restrict__ data, const double* __restrict__ traj, const double* __restrict__ Xb,
                               const double* __restrict__ Pb_in, double* __restrict__ pnub, double* __restrict__ theta_aux) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= st.n_act) return;
  ukf_update_bwd_body<d>(st, a, *Cp, active, smodel, pnu, thoff, starts, data, traj, Xb, Pb_in, pnub, theta_aux);
}

// update of step k followed, for the series that go on, by the sigma points of step k + 1: a thread only ever touches its own
// series' state, so the two halves need no grid-wide ordering
template <int d>
__global__ void ukf_fwd_fused(UkfStep st, UkfStep nx, const int* __restrict__ next_idx, const UkfConst* __restrict__ Cp,
                              const int* __restrict__ active, const int* __restrict__ smodel, const double* __restrict__ pnu,
                              const long long* __restrict__ thoff, const long long* __restrict__ starts, const double* __restrict__ data,
                              const double* __restrict__ traj, double* X, double* P, double* __restrict__ ll, double* __restrict__ states,
                              double* __restrict__ covs, double* __restrict__ xprev, double* __restrict__ Pprev, double* __restrict__ Usave,
                              double* __restrict__ theta_aux) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= st.n_act) return;
  ukf_update_fwd_body<d>(st, a, *Cp, active, smodel, pnu, starts, data, traj, X, P, ll, states, covs);
  const int an = next_idx[st.act0 + a];
  if (an >= 0) ukf_sigma_fwd_body<d>(nx, an, *Cp, active, thoff, X, P, xprev, Pprev, Usave, theta_aux);
}

// adjoint of the sigma points of step k + 1 (for the series that had one) followed by the adjoint of the update of step k
template <int d>
__global__ void ukf_bwd_fused(UkfStep st, UkfStep nx, const int* __restrict__ next_idx, const UkfConst* __restrict__ Cp,
                              const int* __restrict__ active, const int* __restrict__ smodel, const double* __restrict__ pnu,
                              const long long* __restrict__ thoff, const double* __restrict__ Usave, const double* __restrict__ grad_aux,
                              double* Xb, double* Pb, const long long* __restrict__ starts, const double* __restrict__ data,
                              const double* __restrict__ traj, double* Pb_in, double* __restrict__ pnub, double* __restrict__ theta_aux) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= st.n_act) return;
  const int an = next_idx[st.act0 + a];
  if (an >= 0) ukf_sigma_bwd_body<d>(nx, an, *Cp, active, thoff, Usave, grad_aux, Xb, Pb);
  ukf_update_bwd_body<d>(st, a, *Cp, active, smodel, pnu, thoff, starts, data, traj, Xb, Pb_in, pnub, theta_aux);
}

// out = [loss (M) | d loss / d process_model (M x n_pm) | d loss / d Lf (M x d*d)]; block m = model m, fixed summation order
__global__ void ukf_finish(int M, const long long* __restrict__ mser0, const int* __restrict__ mblk_off, const int* __restrict__ mblk,
                           const UkfBlock* __restrict__ blk, const UkfConst* __restrict__ Cp, const double* __restrict__ ll,
                           const double* __restrict__ pnub, const double* __restrict__ grad_aux, const double* __restrict__ pm_all,
                           long long n_pm, long long off_w1, long long n_w1, long long off_w2, long long n_w2,
                           const double* __restrict__ Lf_all, int want_grad, double* __restrict__ out) {
  const UkfConst& C = *Cp;
  const int d = C.d;
  const double reg = C.reg;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const long long s0 = mser0[m], s1 = mser0[m + 1];
    const double* pm = pm_all + (long long)m * n_pm;
    const double* Lf = Lf_all + (long long)m * d * d;
    if (threadIdx.x == 0) {
      double L = 0.0;
      for (long long s = s0; s < s1; ++s) L -= ll[s];
      double r = 0.0;
      for (long long i = 0; i < n_w1; ++i) r = fma(pm[off_w1 + i], pm[off_w1 + i], r);
      for (long long i = 0; i < n_w2; ++i) r = fma(pm[off_w2 + i], pm[off_w2 + i], r);
      out[m] = L + reg * r;
    }
    if (!want_grad) continue;
    double* g_pm = out + M + (long long)m * n_pm;
    double* g_lf = out + M + (long long)M * n_pm + (long long)m * d * d;
    for (long long k = threadIdx.x; k < n_pm + d * d; k += blockDim.x) {
      if (k < n_pm) {
        double g = 0.0;
        for (int q = mblk_off[m]; q < mblk_off[m + 1]; ++q) g += grad_aux[blk[mblk[q]].pm_off + k];
        const bool isw = (k >= off_w1 && k < off_w1 + n_w1) || (k >= off_w2 && k < off_w2 + n_w2);
        if (isw) g = fma(2.0 * reg, pm[k], g);
        g_pm[k] = g;
      } else {                       // Lf_bar = (Pnu_bar + Pnu_bar') Lf
        const int e = (int)(k - n_pm), i = e % d, j = e / d;
        double v = 0.0;
        for (int q = 0; q < d; ++q) {
          double pb = 0.0;
          for (long long s = s0; s < s1; ++s) pb += pnub[s * d * d + i + d * q] + pnub[s * d * d + q + d * i];
          v += pb * Lf[q + d * j];
        }
        g_lf[e] = v;
      }
    }
  }
}

#define UKF_DISPATCH(dd, KERNEL, ...)                                                        \
  switch (dd) {                                                                              \
    case 1: KERNEL<1> __VA_ARGS__; break; case 2: KERNEL<2> __VA_ARGS__; break;              \
    case 3: KERNEL<3> __VA_ARGS__; break; case 4: KERNEL<4> __VA_ARGS__; break;              \
    case 5: KERNEL<5> __VA_ARGS__; break; case 6: KERNEL<6> __VA_ARGS__; break;              \
    case 7: KERNEL<7> __VA_ARGS__; break; default: KERNEL<8> __VA_ARGS__; break;             \
  }

int ukf_build(ude_handle* h, ude_ukf_plan*& out_plan) {
  const ude_desc& D = h->desc;
  if (D.has_series_param) return fail(h, UDE_ERR_UNSUPPORTED, "UKF loss: per-series parameter vectors are not supported");
  if (D.dtype != UDE_F64) return fail(h, UDE_ERR_UNSUPPORTED, "UKF loss: FP64 only");
  if (D.n_cov > 0 && !h->have_cov) return fail(h, UDE_ERR_STATE, "ude_set_covariates has not been called");
  const int d = D.d, n_sig = 2 * d + 1;
  const int64_t M = h->n_models;
  std::unique_ptr<ude_ukf_plan> pl(new ude_ukf_plan());
  pl->d = d; pl->M = M;
  int64_t max_len = 0;
  for (int64_t s = 0; s < h->n_series; ++s) max_len = std::max(max_len, h->lengths[s]);
  const int K = (int)(max_len - 1);
  if (K < 1) return fail(h, UDE_ERR_INVALID, "UKF loss: every series has a single point");
  for (int64_t m = 0; m < M; ++m) {
    int64_t ml = 0;
    for (int64_t s = h->model_series0[(size_t)m]; s < h->model_series0[(size_t)m + 1]; ++s) ml = std::max(ml, h->lengths[(size_t)s]);
    if (ml < 2) return fail(h, UDE_ERR_INVALID, "UKF loss: a model whose series all have a single point");
  }
  pl->K = K;
  // host copies of what the auxiliary description is made from
  UDE_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  std::vector<double> times((size_t)h->n_cols);
  UDE_CUDA_TRY(h, cudaMemcpy(times.data(), h->d_times.p, times.size() * sizeof(double), cudaMemcpyDeviceToHost));
  std::vector<long long> koff; std::vector<double> kt, kx;
  if (D.n_cov > 0) {
    koff.resize((size_t)h->n_series * D.n_cov + 1);
    UDE_CUDA_TRY(h, cudaMemcpy(koff.data(), h->d_knot_off.p, koff.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    kt.resize((size_t)std::max<long long>(koff.back(), 1)); kx.resize(kt.size());
    if (koff.back() > 0) {
      UDE_CUDA_TRY(h, cudaMemcpy(kt.data(), h->d_knot_t.p, (size_t)koff.back() * sizeof(double), cudaMemcpyDeviceToHost));
      UDE_CUDA_TRY(h, cudaMemcpy(kx.data(), h->d_knot_x.p, (size_t)koff.back() * sizeof(double), cudaMemcpyDeviceToHost));
    }
  }
  std::vector<int> active, smodel((size_t)h->n_series);
  for (int64_t s = 0; s < h->n_series; ++s) smodel[(size_t)s] = (int)h->model_of_series[(size_t)s];
  std::vector<long long> thoff;
  std::vector<int64_t> a_starts, a_lengths, a_model, a_koff(1, 0);
  std::vector<double> a_times, a_kt, a_kx;
  int64_t col = 0, tbase = 0;
  const int64_t cols_per_series = 2 * n_sig;
  for (int k = 0; k < K; ++k) {
    UkfStep st; st.k = k; st.act0 = (long long)active.size(); st.col0 = col; st.n_act = 0;
    st.blk0 = (int)pl->blocks.size(); st.n_blk = 0;
    int cur_model = -1;
    for (int64_t s = 0; s < h->n_series; ++s) {             // series order = model order (the models' series are contiguous)
      if (h->lengths[s] <= k + 1) continue;
      const int m = smodel[(size_t)s];
      if (m != cur_model) {                                  // close the previous auxiliary model, open the next
        if (cur_model >= 0) { pl->blocks.back().pm_off = tbase; tbase += D.n_pm; }
        UkfBlock b; b.base = tbase; b.pm_off = 0; b.model = m;
        pl->blocks.push_back(b); ++st.n_blk; cur_model = m;
      }
      active.push_back((int)s); ++st.n_act;
      thoff.push_back(tbase);
      tbase += (int64_t)d * cols_per_series;
      // Reference window: ukf_update receives (times[t], dt) and predict integrates over [times[t], times[t] + dt]
      // (src/UnscentedKalmanFilter.jl:67, :83, :93, :121) -- it STARTS at the end of the observation interval.  Same flow
      // for an autonomous RHS, different with covariates / a time input.  UDE_OPT_UKF_CAUSAL_INTERVAL opts into
      // [times[t-1], times[t]].
      const double t_lo = times[(size_t)(h->starts[s] + k)], t_hi = times[(size_t)(h->starts[s] + k + 1)];
      const double ta = h->ukf_causal ? t_lo : t_hi, tb = h->ukf_causal ? t_hi : t_hi + (t_hi - t_lo);
      for (int j = 0; j < n_sig; ++j) {
        a_starts.push_back(col); a_lengths.push_back(2); a_model.push_back((int64_t)pl->blocks.size() - 1);
        a_times.push_back(ta); a_times.push_back(tb);
        col += 2;
        for (int v = 0; v < D.n_cov; ++v) {       // the knots that bracket [ta, tb] (flat extrapolation keeps the end knots)
          const long long k0 = koff[(size_t)s * D.n_cov + v], k1 = koff[(size_t)s * D.n_cov + v + 1];
          if (k1 == k0) { a_koff.push_back((int64_t)a_kt.size()); continue; }
          long long lo = k0, hi = k1;              // [lo, hi)
          while (lo + 1 < k1 && kt[(size_t)lo + 1] <= ta) ++lo;
          hi = lo + 1;
          while (hi < k1 && kt[(size_t)hi - 1] < tb) ++hi;
          for (long long q = lo; q < hi; ++q) { a_kt.push_back(kt[(size_t)q]); a_kx.push_back(kx[(size_t)q]); }
          a_koff.push_back((int64_t)a_kt.size());
        }
      }
    }
    if (st.n_act == 0) return fail(h, UDE_ERR_INVALID, "UKF loss: internal step without series");
    pl->blocks.back().pm_off = tbase; tbase += D.n_pm;
    pl->steps.push_back(st);
  }
  const int64_t NB = (int64_t)pl->blocks.size();
  pl->NB = NB; pl->n_aux_cols = col; pl->n_aux_theta = tbase;
  // the auxiliary handle: state-space loss r' I r on one-interval mini-series, one model per (step, model)
  ude_desc A = D;
  A.loss_kind = UDE_LOSS_STATE_SPACE; A.proc_inv_dt = 0; A.preroll_eps = 0.0;
  A.lanes_per_segment = 0; A.hidden_per_lane = 0; A.weights_in_smem = 0;
  for (int i = 0; i < UDE_MAX_D * UDE_MAX_D; ++i) { A.A[i] = 0.0; A.B[i] = 0.0; }
  for (int i = 0; i < d; ++i) A.B[i + d * i] = 1.0;
  int rc = ude_create(&A, &pl->aux);
  if (rc != UDE_OK) return fail(h, rc, std::string("UKF auxiliary handle: ") + ude_last_error(nullptr));
  auto aux_fail = [&](int code) { return fail(h, code, std::string("UKF auxiliary handle: ") + ude_last_error(pl->aux)); };
  std::vector<double> zeros((size_t)d * (size_t)col, 0.0);
  rc = ude_set_data(pl->aux, (int64_t)a_starts.size(), a_starts.data(), a_lengths.data(), a_model.data(), col, a_times.data(), zeros.data());
  if (rc != UDE_OK) return aux_fail(rc);
  std::vector<double> ones((size_t)NB, 1.0), nulls((size_t)NB, 0.0);
  rc = ude_set_model_weights(pl->aux, nulls.data(), ones.data(), nulls.data(), nulls.data());
  if (rc != UDE_OK) return aux_fail(rc);
  if (D.n_cov > 0) {
    if (a_kt.empty()) { a_kt.push_back(0.0); a_kx.push_back(0.0); }
    rc = ude_set_covariates(pl->aux, a_koff.data(), a_kt.data(), a_kx.data());
    if (rc != UDE_OK) return aux_fail(rc);
  }
  if (pl->aux->n_theta != pl->n_aux_theta) return fail(h, UDE_ERR_INVALID, "UKF loss: auxiliary layout mismatch");
  rc = build_plan(pl->aux);
  if (rc != UDE_OK) return aux_fail(rc);
  if ((int64_t)pl->aux->model_task0.size() != NB + 1) return fail(h, UDE_ERR_INVALID, "UKF loss: auxiliary task map missing");
  for (int64_t b = 0; b < NB; ++b)
    if (pl->aux->theta_base[(size_t)b] != pl->blocks[(size_t)b].base) return fail(h, UDE_ERR_INVALID, "UKF loss: auxiliary layout mismatch");
  cudaStream_t s = h->stream;
  const size_t NA = active.size(), NS = (size_t)h->n_series;
  std::vector<int> next_idx(active.size(), -1);
  for (int k = 0; k + 1 < K; ++k) {
    const UkfStep& a0 = pl->steps[(size_t)k]; const UkfStep& a1 = pl->steps[(size_t)k + 1];
    int j = 0;                                   // both lists are in series order
    for (int a = 0; a < a0.n_act; ++a) {
      while (j < a1.n_act && active[(size_t)a1.act0 + j] < active[(size_t)a0.act0 + a]) ++j;
      if (j < a1.n_act && active[(size_t)a1.act0 + j] == active[(size_t)a0.act0 + a]) next_idx[(size_t)a0.act0 + a] = j;
    }
  }
  // the auxiliary models of each model, in step order (the order the process-model gradients are summed in)
  std::vector<int> mblk_off((size_t)M + 1, 0), mblk((size_t)NB);
  for (int64_t b = 0; b < NB; ++b) ++mblk_off[(size_t)pl->blocks[(size_t)b].model + 1];
  for (int64_t m = 0; m < M; ++m) mblk_off[(size_t)m + 1] += mblk_off[(size_t)m];
  {
    std::vector<int> fill(mblk_off.begin(), mblk_off.end() - 1);
    for (int64_t b = 0; b < NB; ++b) mblk[(size_t)fill[(size_t)pl->blocks[(size_t)b].model]++] = (int)b;
  }
  UDE_CUDA_TRY(h, pl->d_active.upload(active.data(), NA, s));
  UDE_CUDA_TRY(h, pl->d_next.upload(next_idx.data(), NA, s));
  UDE_CUDA_TRY(h, pl->d_thoff.upload(thoff.data(), NA, s));
  UDE_CUDA_TRY(h, pl->d_smodel.upload(smodel.data(), NS, s));
  UDE_CUDA_TRY(h, pl->d_blk.upload(pl->blocks.data(), (size_t)NB, s));
  UDE_CUDA_TRY(h, pl->d_mblk_off.upload(mblk_off.data(), mblk_off.size(), s));
  UDE_CUDA_TRY(h, pl->d_mblk.upload(mblk.data(), mblk.size(), s));
  UDE_CUDA_TRY(h, pl->d_xprev.reserve(NA * d)); UDE_CUDA_TRY(h, pl->d_Pprev.reserve(NA * d * d)); UDE_CUDA_TRY(h, pl->d_U.reserve(NA * d * d));
  UDE_CUDA_TRY(h, pl->d_x.reserve(NS * d)); UDE_CUDA_TRY(h, pl->d_P.reserve(NS * d * d));
  UDE_CUDA_TRY(h, pl->d_xb.reserve(NS * d)); UDE_CUDA_TRY(h, pl->d_Pb.reserve(NS * d * d));
  UDE_CUDA_TRY(h, pl->d_ll.reserve(NS)); UDE_CUDA_TRY(h, pl->d_pnub.reserve(NS * d * d));
  UDE_CUDA_TRY(h, pl->d_theta_aux.reserve((size_t)pl->n_aux_theta));
  UDE_CUDA_TRY(h, pl->d_grad_aux.reserve((size_t)pl->n_aux_theta));
  UDE_CUDA_TRY(h, pl->d_traj_aux.reserve((size_t)d * (size_t)col));
  UDE_CUDA_TRY(h, pl->d_pm.reserve((size_t)M * ((size_t)D.n_pm + (size_t)d * d)));
  UDE_CUDA_TRY(h, pl->d_pnu.reserve((size_t)M * d * d));
  UDE_CUDA_TRY(h, pl->d_out.reserve((size_t)M * (1 + (size_t)D.n_pm + (size_t)d * d)));
  UDE_CUDA_TRY(h, pl->d_const.reserve(1));
  UDE_CUDA_TRY(h, cudaMemsetAsync(pl->d_theta_aux.p, 0, (size_t)pl->n_aux_theta * sizeof(double), s));
  UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
  out_plan = pl.release();
  return UDE_OK;
}

// every launch of one evaluation, in order, on `s` (capturable: no host-dependent arguments)
int ukf_enqueue(ude_handle* h, ude_ukf_plan* pl, bool want_grad, double* d_states, double* d_covs, cudaStream_t s) {
  const ude_desc& D = h->desc;
  const int d = D.d, K = pl->K;
  const int M = (int)pl->M, NB = (int)pl->NB;
  ude_handle* ax = pl->aux;
  const size_t NS = (size_t)h->n_series;
  const UkfConst* C = pl->d_const.p;
  const int* act = pl->d_active.p; const int* smod = pl->d_smodel.p;
  const double* pnu = pl->d_pnu.p; const long long* thoff = pl->d_thoff.p;
  const int TB = 64;
  auto task0 = [&](const UkfStep& st) { return ax->model_task0[(size_t)st.blk0]; };
  auto task1 = [&](const UkfStep& st) { return ax->model_task0[(size_t)st.blk0 + (size_t)st.n_blk]; };
  UDE_CUDA_TRY(h, cudaMemsetAsync(pl->d_ll.p, 0, NS * sizeof(double), s));
  ukf_init<<<(int)((NS + TB - 1) / TB), TB, 0, s>>>((int)NS, d, h->d_starts.p, h->d_data.p, C, pl->d_x.p, pl->d_P.p, d_states, d_covs);
  ukf_bcast_pm<<<NB, 128, 0, s>>>(pl->d_pm.p, D.n_pm, pl->d_blk.p, NB, pl->d_theta_aux.p);
  UDE_CUDA_TRY(h, cudaGetLastError());
  // per step: [sigma points] -> segment kernel (2d+1 predicts per series) -> [update]; the update of step k and the sigma
  // points of step k+1 are one launch
  {
    const UkfStep& st0 = pl->steps[0];
    UKF_DISPATCH(d, ukf_sigma_fwd, <<<(st0.n_act + TB - 1) / TB, TB, 0, s>>>(st0, C, act, thoff, pl->d_x.p, pl->d_P.p, pl->d_xprev.p, pl->d_Pprev.p,
                                                                              pl->d_U.p, pl->d_theta_aux.p))
  }
  for (int k = 0; k < K; ++k) {
    const UkfStep& st = pl->steps[(size_t)k];
    const int nb = (st.n_act + TB - 1) / TB;
    int rc = launch_segments(ax, pl->d_theta_aux.p, nullptr, pl->d_traj_aux.p, task0(st), task1(st), 0, s);
    if (rc != UDE_OK) return fail(h, rc, std::string("UKF predict: ") + ude_last_error(ax));
    if (k + 1 < K) {
      const UkfStep& nx = pl->steps[(size_t)k + 1];
      UKF_DISPATCH(d, ukf_fwd_fused, <<<nb, TB, 0, s>>>(st, nx, pl->d_next.p, C, act, smod, pnu, thoff, h->d_starts.p, h->d_data.p, pl->d_traj_aux.p,
                                                       pl->d_x.p, pl->d_P.p, pl->d_ll.p, d_states, d_covs, pl->d_xprev.p, pl->d_Pprev.p, pl->d_U.p,
                                                       pl->d_theta_aux.p))
    } else {
      UKF_DISPATCH(d, ukf_update_fwd, <<<nb, TB, 0, s>>>(st, C, act, smod, pnu, h->d_starts.p, h->d_data.p, pl->d_traj_aux.p, pl->d_x.p,
                                                        pl->d_P.p, pl->d_ll.p, d_states, d_covs))
    }
  }
  UDE_CUDA_TRY(h, cudaGetLastError());
  if (want_grad) {
    UDE_CUDA_TRY(h, cudaMemsetAsync(pl->d_grad_aux.p, 0, (size_t)pl->n_aux_theta * sizeof(double), s));
    UDE_CUDA_TRY(h, cudaMemsetAsync(pl->d_xb.p, 0, NS * d * sizeof(double), s));
    UDE_CUDA_TRY(h, cudaMemsetAsync(pl->d_Pb.p, 0, NS * d * d * sizeof(double), s));
    UDE_CUDA_TRY(h, cudaMemsetAsync(pl->d_pnub.p, 0, NS * d * d * sizeof(double), s));
    UDE_CUDA_TRY(h, cudaMemsetAsync(ax->d_model_loss.p, 0, (size_t)NB * sizeof(double), s));
    // reverse: [update adjoint] -> segment loss+adjoint kernel -> [sigma adjoint]; the sigma adjoint of step k+1 and the update
    // adjoint of step k are one launch; the sigma adjoint of step 0 would differentiate w.r.t. the fixed initial state: skipped
    {
      const UkfStep& stl = pl->steps[(size_t)K - 1];
      UKF_DISPATCH(d, ukf_update_bwd, <<<(stl.n_act + TB - 1) / TB, TB, 0, s>>>(stl, C, act, smod, pnu, thoff, h->d_starts.p, h->d_data.p, pl->d_traj_aux.p,
                                                                                 pl->d_xb.p, pl->d_Pb.p, pl->d_pnub.p, pl->d_theta_aux.p))
    }
    for (int k = K - 1; k >= 0; --k) {
      const UkfStep& sk = pl->steps[(size_t)k];
      int rc = launch_segments(ax, pl->d_theta_aux.p, pl->d_grad_aux.p, nullptr, task0(sk), task1(sk), 0, s);
      if (rc != UDE_OK) return fail(h, rc, std::string("UKF adjoint: ") + ude_last_error(ax));
      if (ax->n_models == 1) {
        // a single auxiliary model (one filter step, one model): the single-model kernels leave the process-model gradient in
        // per-block partial rows, which only the finalize kernel folds into the gradient vector
        rc = finalize_eval(ax, pl->d_theta_aux.p, ax->d_loss.p, pl->d_grad_aux.p, ax->grid_grad, s);
        if (rc != UDE_OK) return fail(h, rc, std::string("UKF adjoint: ") + ude_last_error(ax));
      }
      if (k > 0) {
        const UkfStep& st = pl->steps[(size_t)k - 1]; const UkfStep& nx = pl->steps[(size_t)k];
        UKF_DISPATCH(d, ukf_bwd_fused, <<<(st.n_act + TB - 1) / TB, TB, 0, s>>>(st, nx, pl->d_next.p, C, act, smod, pnu, thoff, pl->d_U.p, pl->d_grad_aux.p,
                                                                               pl->d_xb.p, pl->d_Pb.p, h->d_starts.p, h->d_data.p, pl->d_traj_aux.p,
                                                                               pl->d_Pb.p, pl->d_pnub.p, pl->d_theta_aux.p))
      }
    }
    UDE_CUDA_TRY(h, cudaGetLastError());
  }
  const long long n_w1 = (long long)D.nn_hidden * D.nn_in, n_w2 = (long long)D.nn_out * D.nn_hidden;
  ukf_finish<<<std::min(M, 1024), 128, 0, s>>>(M, h->d_model_series0.p, pl->d_mblk_off.p, pl->d_mblk.p, pl->d_blk.p, C, pl->d_ll.p, pl->d_pnub.p,
                                               pl->d_grad_aux.p, pl->d_pm.p, D.n_pm, D.off_w1, n_w1, D.off_w2, n_w2,
                                               pl->d_pm.p + (size_t)M * D.n_pm, want_grad ? 1 : 0, pl->d_out.p);
  UDE_CUDA_TRY(h, cudaGetLastError());
  return UDE_OK;
}

// One evaluation of every model of the handle.  theta: host, n_theta ([uhat | process model] per model); Lf: host, d*d
// column-major per model (P_nu = Lf Lf'); loss: n_models; grad_lf: n_models * d*d.  states/covs (host, may be NULL): filtered
// means d x n_cols and covariances d x d x n_cols (ukf_smoothing).  The launch sequence (2 kernels per filter step and
// direction) is captured once per variant and replayed as a CUDA graph: a single series of 60 points is launch-bound otherwise.
int ukf_eval(ude_handle* h, const ude_ukf_opts* o, const double* theta, const double* Lf, double* loss, double* grad_theta,
             double* grad_lf, double* states, double* covs) {
  const ude_desc& D = h->desc;
  if (!h->ukf) { int rc = ukf_build(h, h->ukf); if (rc != UDE_OK) return rc; }
  ude_ukf_plan* pl = h->ukf;
  const int d = D.d, n_sig = 2 * d + 1;
  const size_t M = (size_t)pl->M, n_pm = (size_t)D.n_pm, dd = (size_t)d * d;
  UkfConst C; memset(&C, 0, sizeof C);
  C.d = d; C.n_sig = n_sig; C.reg = o->reg;
  const double lam = o->alpha * o->alpha * (d - o->kappa) - d;
  C.c = lam + d; C.Wm0 = lam / (lam + d); C.Wc0 = lam / (lam + d) + (1.0 - o->alpha * o->alpha + o->beta);
  C.Wi = 1.0 / (2.0 * (d + lam)); C.ll_const = 0.5 * d * log(2.0 * 3.14159);      // sic: the reference's constant
  for (size_t i = 0; i < dd; ++i) C.Peta[i] = o->Peta[i];
  // host staging: [pm of every model | Lf of every model], and P_nu = Lf Lf' of every model
  std::vector<double> stage(M * (n_pm + dd)), pnu(M * dd);
  auto pm_of = [&](size_t m) -> const double* { return theta + h->theta_base[m] + (int64_t)d * (h->model_col0[m + 1] - h->model_col0[m]); };
  for (size_t m = 0; m < M; ++m) {
    memcpy(stage.data() + m * n_pm, pm_of(m), n_pm * sizeof(double));
    const double* L = Lf + m * dd;
    memcpy(stage.data() + M * n_pm + m * dd, L, dd * sizeof(double));
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) {
        double v = 0.0;
        for (int k = 0; k < d; ++k) v += L[i + d * k] * L[j + d * k];
        pnu[m * dd + i + (size_t)d * j] = v;
      }
  }
  const bool want_grad = grad_theta != nullptr || grad_lf != nullptr;
  const bool want_states = states != nullptr;
  cudaStream_t s = h->stream;
  if (want_states) {
    UDE_CUDA_TRY(h, pl->d_states.reserve((size_t)d * h->n_cols));
    if (covs) UDE_CUDA_TRY(h, pl->d_covs.reserve((size_t)d * d * h->n_cols));
  }
  double* d_states = want_states ? pl->d_states.p : nullptr;
  double* d_covs = (want_states && covs) ? pl->d_covs.p : nullptr;
  // pageable sources that live on this stack frame: synchronous copies before anything is enqueued
  UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
  UDE_CUDA_TRY(h, cudaMemcpy(pl->d_const.p, &C, sizeof C, cudaMemcpyHostToDevice));
  UDE_CUDA_TRY(h, cudaMemcpy(pl->d_pm.p, stage.data(), stage.size() * sizeof(double), cudaMemcpyHostToDevice));
  UDE_CUDA_TRY(h, cudaMemcpy(pl->d_pnu.p, pnu.data(), pnu.size() * sizeof(double), cudaMemcpyHostToDevice));
  const int variant = want_grad ? 1 : (want_states ? (covs ? 3 : 2) : 0);
  const char* genv = getenv("UDE_UKF_GRAPH");
  pl->used_graph = 0;
  if (!(genv && atoi(genv) == 0)) {
    if (!pl->exec[variant]) {
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int rc = ukf_enqueue(h, pl, want_grad, d_states, d_covs, s);
        const cudaError_t e = cudaStreamEndCapture(s, &g);
        if (rc == UDE_OK && e == cudaSuccess && g) {
          if (cudaGraphInstantiate(&pl->exec[variant], g, 0) != cudaSuccess) pl->exec[variant] = nullptr;
        }
        if (g) cudaGraphDestroy(g);
      }
      cudaGetLastError();
    }
    if (pl->exec[variant]) { UDE_CUDA_TRY(h, cudaGraphLaunch(pl->exec[variant], s)); pl->used_graph = 1; }
  }
  if (!pl->used_graph) { int rc = ukf_enqueue(h, pl, want_grad, d_states, d_covs, s); if (rc != UDE_OK) return rc; }
  std::vector<double> outv(M * (1 + n_pm + dd), 0.0);
  UDE_CUDA_TRY(h, cudaMemcpyAsync(outv.data(), pl->d_out.p, (want_grad ? outv.size() : M) * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (want_states) {
    UDE_CUDA_TRY(h, cudaMemcpyAsync(states, d_states, (size_t)d * h->n_cols * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (covs) UDE_CUDA_TRY(h, cudaMemcpyAsync(covs, d_covs, (size_t)d * d * h->n_cols * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  UDE_CUDA_TRY(h, cudaStreamSynchronize(s));
  bool finite = true;
  for (size_t m = 0; m < M; ++m) { if (loss) loss[m] = outv[m]; finite = finite && std::isfinite(outv[m]); }
  if (grad_theta) {
    memset(grad_theta, 0, (size_t)h->n_theta * sizeof(double));      // uhat does not enter this loss
    for (size_t m = 0; m < M; ++m) memcpy(grad_theta + (pm_of(m) - theta), outv.data() + M + m * n_pm, n_pm * sizeof(double));
  }
  if (grad_lf) memcpy(grad_lf, outv.data() + M + M * n_pm, M * dd * sizeof(double));
  if (!finite) return fail(h, UDE_NONFINITE, "UKF loss is not finite (covariance lost positive definiteness?)");
  return UDE_OK;
}

}  // namespace
