/* ude_cuda.h -- C ABI of libude_cuda.so, the B200 (sm_100a) loss-and-gradient engine for
 * UniversalDiffEq-style universal differential equations.
 *
 * The reference (Jack-H-Buckner/UniversalDiffEq.jl) has NO foreign-function boundary on this path: the
 * loss is a Julia closure  loss(theta::ComponentVector)::Float64  built by
 *     conditional_likelihood / multiple_shooting_loss / shooting_loss / derivative_matching_loss
 *     (src/LossFunctionConstructors.jl:17-98, :261-305, :639-708, :196-232, :545-601, :175-192, :500-540)
 *     init_loss / init_multi_loss_function (src/helpers.jl:16-76, src/MultipleTimeSeries.jl:121-157)
 * and differentiated by Zygote through  Optimization.OptimizationFunction(target, AutoZygote())
 * (src/Optimizers.jl:473-478, :658-663).  The entry points below are what a `ccall` from those two places
 * binds instead (INTEGRATION.md shows the Julia stub): the handle captures what the closure captured
 * (data, times, covariates, loss options, solver), and ude_loss_grad is the closure + its pullback.
 *
 * Conventions: plain C, host pointers unless the name ends in _device, FP64 arrays, column-major matrices
 * exactly as Julia stores them, 0-based series offsets (the Julia stub subtracts 1 from `starts`,
 * src/helpers.jl:516-526).  Every call returns UDE_OK (0) or a negative ude_status; the message is
 * available from ude_last_error.  Nothing throws across the boundary.  Handles are independent: different
 * handles may be used from different host threads concurrently (src/CrossValidation.jl:30, :259 call the
 * optimiser from Threads.@threads); one handle is not re-entrant.
 */
#ifndef UDE_CUDA_H
#define UDE_CUDA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UDE_ABI_VERSION 1
#define UDE_MAX_SCALARS 8
#define UDE_MAX_D 8

typedef enum ude_status {
  UDE_OK = 0,
  UDE_ERR_INVALID = -1,      /* bad argument / inconsistent description */
  UDE_ERR_UNSUPPORTED = -2,  /* shape or RHS kind with no compiled sm_100a kernel */
  UDE_ERR_CUDA = -3,         /* CUDA runtime error (message has the cudaError string) */
  UDE_ERR_NO_DEVICE = -4,    /* no usable sm_100 device: there is NO CPU fallback */
  UDE_ERR_STATE = -5,        /* call order (e.g. loss_grad before set_data) */
  UDE_ERR_NCCL = -6,
  UDE_NONFINITE = -7         /* the LOSS is not finite (checked on the host after the call; the gradient is not scanned);
                                outputs are still written (loss = Inf/NaN as computed) */
} ude_status;

/* solver: fixed-step explicit RK replacing OrdinaryDiffEq.solve(..., Tsit5(), abstol=1e-6, reltol=1e-6)
 * (src/ProcessModels.jl:82-83); `substeps` steps per observation interval. */
enum { UDE_RK4 = 0, UDE_TSIT5 = 1,
       UDE_EULER = 2 /* one RHS evaluation per step: the spline-gradient-matching loss (LFC.jl:888-931) is a state-space
                        loss on the knot grid with this solver; FP64 lane-per-hidden-unit kernels, state-space loss only */ };
enum { UDE_F64 = 0, UDE_F32 = 1 };
/* loss kinds */
enum {
  UDE_LOSS_STATE_SPACE = 0, /* conditional likelihood (LFC.jl:17-98, :347-404) and the default MSE loss (helpers.jl:16-76) */
  UDE_LOSS_MULTI_SHOOT = 1, /* LFC.jl:261-305, :639-708 */
  UDE_LOSS_SHOOT = 2,       /* LFC.jl:196-232, :545-601 */
  UDE_LOSS_DERIV_MATCH = 3  /* LFC.jl:175-192, :500-540 */
};
/* right-hand-side catalog: the shapes users of NODE / MultiNODE / CustomDerivatives write */
enum {
  UDE_RHS_NODE = 0,          /* du = NN([u; X(t)])                      src/ProcessModels.jl:387-390, :434-437 */
  UDE_RHS_NODE_TIME = 1,     /* du[1:nn_out] = NN([u; X(t); a t + b]); du[j>nn_out] = -s_j u_j   docs/src/examples.md:44-49 */
  UDE_RHS_LV_UDE = 2,        /* du1 = r u1 - NN(u); du2 = theta NN(u) - m u2 (+ beta_i X1)        test/UDETests.jl:20-28, :82-86 */
  UDE_RHS_LV_UDE_ABS = 3,    /* README.md:52-62 (abs on NN and parameters, logistic prey growth) */
  UDE_RHS_FISHERY = 4,       /* docs/src/examples.md:163-166 */
  UDE_RHS_NN_DECAY = 5,      /* du = NN(u) - m u                         test/MultiUDE.jl:17-19 */
  UDE_RHS_SERIES_GROWTH = 6, /* du = r[series] u NN(u)[1]                docs/src/MultipleTimeSeries.md:51-55 */
  /* >= 1000: a user-written right-hand side (the reference's `derivs!` closure, src/Models.jl:115-116).  Its kernels are not in
   * this library: the host side writes the parametric part as expressions, generates the input / combine / combine_vjp trait
   * of csrc/ude_kernel_body.inc from them, compiles it with nvcc into a PLUGIN shared library linked against libude_cuda.so
   * and dlopens it (RTLD_GLOBAL) -- the plugin's static initialiser registers its kernels under this rhs_kind, after which
   * every entry point below accepts it (universaldiffeq.jl_b200/user_rhs.py). */
  UDE_RHS_USER0 = 1000
};

typedef struct ude_handle ude_handle;

/* Everything the reference's loss closure captures that is not an array. Offsets are in doubles, relative
 * to the start of `process_model` inside one model's theta = [uhat (d x n_cols_of_model, column-major) |
 * process_model | nothing else] (src/helpers.jl:3-11; the other ComponentArray groups are empty). */
typedef struct ude_desc {
  int32_t abi_version;        /* UDE_ABI_VERSION */
  int32_t struct_bytes;       /* sizeof(ude_desc) */
  int32_t device;             /* CUDA ordinal */
  int32_t dtype;              /* UDE_F64 (UDE_F32: arithmetic in FP32, reductions in FP64) */
  int32_t solver, substeps;
  int32_t rhs_kind, d, n_cov, nn_in, nn_hidden, nn_out, n_scalars, has_series_param;
  int64_t off_w1, off_b1, off_w2, off_b2;           /* Lux Dense leaves: weight (out x in, column-major), bias */
  int64_t off_scalar[UDE_MAX_SCALARS];
  int64_t off_series_param;                         /* vector indexed by the series' position inside its model */
  int64_t n_pm;                                     /* length of process_model */
  double time_scale, time_shift;                    /* UDE_RHS_NODE_TIME input feature a*t + b */
  int32_t loss_kind, pred_length, proc_inv_dt, shoot_from_theta, shoot_first_scored, remove_ends;
  double preroll_eps;                               /* LFC.jl:268: tspan starts 1e-6 before the block (single-series only) */
  double A[UDE_MAX_D * UDE_MAX_D];                  /* observation weight, d x d column-major (Sigma_obs^-1) */
  double B[UDE_MAX_D * UDE_MAX_D];                  /* process weight (Sigma_proc^-1) */
  int32_t lanes_per_segment;  /* 0 = choose; else force the sub-warp group size G (tuning / tests) */
  int32_t hidden_per_lane;    /* 0 = choose */
  int32_t weights_in_smem;    /* 0 = choose; 1 = force register-resident weights; 2 = force shared-memory weights */
  int32_t reserved[5];
} ude_desc;

int ude_abi_version(void);
int ude_device_count(void);

/* Allocates a handle bound to desc->device. Fails with UDE_ERR_NO_DEVICE when no GPU is present. */
int ude_create(const ude_desc* desc, ude_handle** out);
int ude_destroy(ude_handle* h);
/* Message of the last failure on this handle (h == NULL: last failure of a create on this thread). */
const char* ude_last_error(const ude_handle* h);

/* data d x n_cols column-major, times n_cols, both sorted by (series, time) as process_multi_data does
 * (src/helpers.jl:529-556); series s owns columns [starts[s], starts[s]+lengths[s]).  model_of_series
 * (NULL = one model) lets one handle hold many independent theta vectors (CV folds x ensemble members,
 * src/CrossValidation.jl:17-45): theta_all = concat over models of [uhat_m | process_model_m]. */
int ude_set_data(ude_handle* h, int64_t n_series, const int64_t* starts, const int64_t* lengths,
                 const int64_t* model_of_series, int64_t n_cols, const double* times, const double* data);
/* per-model loss weights, each n_models long (NULL = ones / zeros for reg_weight):
 *   obs_scale*A and proc_scale*B(/dt if proc_inv_dt) weight the state-space loss (src/LossFunctions.jl:17-28),
 *   shoot_weight scales the shooting loss, reg_weight*(|W1|^2+|W2|^2) is added (src/Regularization.jl:97-105). */
int ude_set_model_weights(ude_handle* h, const double* obs_scale, const double* proc_scale,
                          const double* shoot_weight, const double* reg_weight);
/* piecewise-linear covariates with flat extrapolation (src/covariates.jl:16-19, :45-48):
 * knots of (series s, variable v) are [knot_off[s*n_cov+v], knot_off[s*n_cov+v+1]) */
int ude_set_covariates(ude_handle* h, const int64_t* knot_off, const double* knot_t, const double* knot_x);
/* derivative matching inputs: smoothed states and their time derivatives (LFC.jl:161-171), d x n_cols */
int ude_set_targets(ude_handle* h, const double* smoothed_states, const double* dudt);
/* skip_mask[c] != 0: the process term of the interval ENDING at column c is left out (t_skip, LFC.jl:82) */
int ude_set_skip(ude_handle* h, const uint8_t* skip_mask);

/* Handle options (value 0 = off, the default).
 *   UDE_OPT_SPARSE_UHAT_IO  multiple-shooting and shooting losses read only the BLOCK-START columns of uhat, and only those
 *     columns have a non-zero gradient (LFC.jl:281-296, :665-680: `uhats[:,t]` at block starts; :196-232: `uhat[:,1]`).
 *     With this option ude_loss_grad, when theta and grad are page-locked host buffers (cudaHostAlloc / cudaHostRegister),
 *     lets the kernels read those columns of theta and write those columns of grad directly over PCIe (zero-copy): only
 *     8*(d*n_blocks + n_pm) bytes move each way instead of 8*n_theta.  CONTRACT: the structurally-zero entries of grad are
 *     NOT written -- the caller zeroes the gradient buffer once and reuses it (every entry the library does not write stays
 *     exactly 0).  Ignored (dense copies, every entry written) for other losses, many-model handles or pageable buffers.
 *   UDE_OPT_UKF_CAUSAL_INTERVAL  the reference's UKF hands (times[t], dt) to predict, i.e. it integrates the sigma points
 *     over [times[t], times[t] + dt] -- the window starts at the END of the observation interval
 *     (src/UnscentedKalmanFilter.jl:67, :83, :93, :121).  Default 0 reproduces that; 1 uses [times[t-1], times[t]].
 *     Identical for autonomous right-hand sides, different with covariates or a time input. */
enum { UDE_OPT_SPARSE_UHAT_IO = 1, UDE_OPT_UKF_CAUSAL_INTERVAL = 2 };
int ude_set_option(ude_handle* h, int32_t option, int64_t value);
/* bytes the last ude_loss_grad call moved host->device and device->host (copies + zero-copy reads/writes) */
int ude_last_io_bytes(const ude_handle* h, int64_t* h2d_bytes, int64_t* d2h_bytes);

int64_t ude_n_theta(const ude_handle* h);
int64_t ude_n_models(const ude_handle* h);
/* RK steps one loss evaluation integrates (unit of the throughput metric) and kernels launched per call */
int64_t ude_steps_per_eval(const ude_handle* h);
/* name of the compiled kernel specialisation the handle dispatches to (e.g. "node_d4x1_mma_h32") */
const char* ude_kernel_name(ude_handle* h);
int32_t ude_launches_per_eval(const ude_handle* h);

/* loss[n_models] and grad[n_theta] (grad may be NULL: loss only). Synchronous, host buffers:
 * H2D(theta) -> kernels (-> allreduce) -> D2H(loss, grad).  Large problems are pipelined inside the call (page-locked
 * buffers make the copies asynchronous): a single-model handle in chunks of series, a many-model handle in runs of whole
 * models (theta of run k+1 up, gradient and losses of run k-1 down, while run k is on the SMs); with
 * UDE_OPT_SPARSE_UHAT_IO the trajectory losses of a single-model handle move only the block-start columns (zero-copy).
 * Tuning aids (environment, read when the plan is built): UDE_HOST_CHUNKS=<n> (1 = no pipeline), UDE_HOST_CHUNK_ROUNDS=0,
 * UDE_HOST_CHUNK_EDGE=0, UDE_KERNEL=<substring of a kernel name>. */
int ude_loss_grad(ude_handle* h, const double* theta, int64_t n_theta, double* loss, double* grad);
/* Same with device-resident theta / loss / grad on `stream` (a cudaStream_t; NULL = the handle's stream).
 * Asynchronous: returns after enqueueing. No nonfinite check. */
int ude_loss_grad_device(ude_handle* h, const double* d_theta, double* d_loss, double* d_grad, void* stream);
/* Forward-only trajectories at every column (shooting_states / multiple_shooting_states / predictions,
 * LFC.jl:237-258, :310-341, :604-631, :711-756; src/ModelTesting.jl:181-195): out is d x n_cols. */
int ude_forward(ude_handle* h, const double* theta, int64_t n_theta, double* out);

/* forecast (src/ModelTesting.jl:541-571): the handle's series are one-interval mini-series [t_{j-1}, t_j] of ONE trajectory;
 * interval j starts from the damped end of interval j-1 (extrapolation damping of init_forecast, src/ProcessModels.jl:2-25:
 * outside [umin, umax] the increment is blended with a pull towards umean, width l, rate rho).  limits = [umax | umin |
 * umean] (3 d); out is d x n_series.  The whole chain stays on the device. */
int ude_forecast_chain(ude_handle* h, const double* theta, int64_t n_theta, const double* u0, const double* limits, double l,
                       double rho, double* out);

/* Multi-GPU (one handle per GPU, one rank per process or per host thread): after ude_comm_init every
 * ude_loss_grad* call sum-allreduces [grad of process_model (minus the per-series vector) | loss] across
 * ranks with one NCCL call on the handle's stream. id_bytes is an ncclUniqueId (128 B). */
int ude_comm_unique_id(void* id_bytes, int32_t n_bytes);
int ude_comm_init(ude_handle* h, const void* id_bytes, int32_t n_bytes, int32_t n_ranks, int32_t rank);

/* Fused exchange over peer memory (NVLink / NVSwitch P2P) -- replaces the NCCL allreduce of ude_comm_init by ONE kernel
 * that finalises the block reduction, pushes [loss | shared process-model gradient] into every rank's exchange buffer
 * with peer stores, and sums the n_ranks contributions in rank order (bit-identical on all ranks).
 *   1. every rank: ude_peer_export(h, n_ranks, blob, 128)   -> allocates the rank's exchange buffer, describes it in blob
 *   2. the host exchanges the blobs (any transport; bench.py uses torch.distributed all_gather)
 *   3. every rank: ude_peer_attach(h, all_blobs, n_ranks, rank)   (all_blobs: n_ranks x 128 B in rank order)
 * Ranks are one process per GPU (CUDA IPC) or threads of one process with one GPU each.  Two ranks on the same GPU are
 * refused.  From here on EVERY loss evaluation on the handle (ude_loss_grad*, ude_adam, loss-only calls) is a collective:
 * all ranks must make the same sequence of calls.  ude_forward is not (it needs no loss) and may be called by one rank
 * alone.  A rank that waits longer than UDE_PEER_TIMEOUT_MS (default 60000) for a peer returns NaN loss/gradient (status
 * UDE_NONFINITE) instead of hanging; that state is STICKY -- the ranks' call counters have diverged, every later evaluation
 * on the handle returns NaN too: destroy the handles and attach again. */
#define UDE_MAX_PEERS 16
#define UDE_PEER_BLOB_BYTES 128
int ude_peer_export(ude_handle* h, int32_t n_ranks, void* blob, int32_t n_bytes);
int ude_peer_attach(ude_handle* h, const void* blobs, int32_t n_ranks, int32_t rank);

/* Device-resident Adam (Optimisers.jl semantics: beta=(0.9,0.999), eps=1e-8, bias-corrected), SURVEY.md 8f row 1.
 * Runs n_iter iterations of loss_grad + update without leaving the device; theta is read and written back.
 * loss_trace (n_iter doubles per model, may be NULL) receives the loss before each update.
 * RETURNED ITERATE: theta after the LAST update (iteration n_iter), not the best-seen one.  The reference's
 * `Optimization.solve(prob, Adam(eta); maxiters)` (src/Optimizers.jl:500, :685) goes through OptimizationOptimisers, which
 * in the versions the compat range allows (0.1 - 0.3, Project.toml:61) returns the iterate with the lowest objective seen
 * when the iteration limit is hit; that is third-party behaviour that cannot be checked here (no Julia).  A caller that
 * wants it takes argmin(loss_trace) and re-runs that many iterations (the iterates are deterministic), as
 * training.train_(..., return_best=True) does.  Returns UDE_NONFINITE (theta still written) when the loss of the last
 * iteration is not finite. */
int ude_adam(ude_handle* h, double* theta, int64_t n_theta, double step_size, int32_t n_iter, double* loss_trace);
/* 1 if the last ude_adam call replayed a captured CUDA graph (one iteration = memset + segment kernel + finalize + update),
 * which is what makes small, launch-bound models train at kernel speed; UDE_ADAM_GRAPH=0 disables it. */
int ude_adam_used_graph(const ude_handle* h);


/* Unscented-Kalman-filter marginal likelihood (src/UnscentedKalmanFilter.jl:2-87; LFC.jl:103-127, :427-476):
 *   loss = - sum_series sum_t loglik_t + reg * (|W1|^2 + |W2|^2),  P_nu = F F' with F = pnu_factor (d x d, column-major),
 * identity observation model, P_eta fixed, first state = first observation (as the reference).  uhat does not enter the
 * loss (its gradient is returned as zeros).  The 2d+1 sigma-point predicts of every step run through the same fused RK
 * kernels as the other losses.  Many-model handles (folds x ensemble members, ude_set_data's
 * model_of_series) are filtered in the same launches: `loss` then holds n_models values, `pnu_factor` / `grad_pnu_factor`
 * n_models blocks of d*d (one process-noise factor per model), alpha / beta / kappa / reg / Peta are shared.  MultiUDE's
 * diagonal form P_nu = I .* p.^2 (src/UnscentedKalmanFilter.jl:123) is F = diag(p).  No per-series parameter vector. */
typedef struct ude_ukf_opts {
  double alpha, beta, kappa;                 /* reference defaults 1e-3, 2, 0 (src/Optimizers.jl:408) */
  double reg;                                /* total weight of (|W1|^2 + |W2|^2) */
  double Peta[UDE_MAX_D * UDE_MAX_D];        /* observation covariance, d x d column-major */
} ude_ukf_opts;
int ude_ukf_loss_grad(ude_handle* h, const ude_ukf_opts* opts, const double* theta, int64_t n_theta, const double* pnu_factor,
                      double* loss /* n_models */, double* grad_theta /* n_theta or NULL */,
                      double* grad_pnu_factor /* n_models * d*d or NULL */);
/* ukf_smoothing: filtered means (d x n_cols) and covariances (d x d x n_cols, may be NULL; zero at a series' first column) */
int ude_ukf_filter(ude_handle* h, const ude_ukf_opts* opts, const double* theta, int64_t n_theta, const double* pnu_factor,
                   double* states, double* covariances);

/* Derivative-matching pre-step, batched (LFC.jl:161-171, :510-522: RegularizationSmooth(u, t, d; alg = :gcv_svd) per state and
 * series + DataInterpolations.derivative at the observation times).  All `rows` rows of U (rows x n, row-major) share one time
 * grid, whose operators the host builds once: Vt (r x n, right singular vectors of the d-th order derivative matrix D), s (r,
 * its singular values), M (n x n: knot derivatives of the natural cubic spline as a matrix).  Per row the kernel minimises
 * GCV(lambda) over [lambda_lo, lambda_hi] (RegularizationTools defaults 1e-4, 1e3), and writes the smoothed values, their
 * time derivative and the chosen lambda.  Host buffers in and out; no handle. */
int ude_reg_smooth(int32_t device, int32_t n, int32_t r, int64_t rows, const double* Vt, const double* s, const double* M,
                   const double* U, double lambda_lo, double lambda_hi, double* out_u, double* out_dudt, double* out_lambda);

/* Kernel timing for roofline reports: after ude_profile_begin(h, n) the next n evaluations bracket the segment
 * kernel (the dominant launch) with CUDA events on the launching stream; ude_profile_read synchronises and returns
 * how many were recorded, writing their durations in milliseconds. */
int ude_profile_begin(ude_handle* h, int32_t max_evals);
int ude_profile_read(ude_handle* h, float* ms_out, int32_t capacity);

/* Measured FP64 FMA throughput of the device (TFLOP/s): the roofline denominator bench.py reports against. */
int ude_measure_fp64_peak(int32_t device, double* tflops);
/* Same for FP32 FMA: the denominator for handles created with dtype = UDE_F32. */
int ude_measure_fp32_peak(int32_t device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
