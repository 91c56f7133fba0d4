/* ude_oracle.h -- CPU ORACLE for the UDE loss-and-gradient hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference (Jack-H-Buckner/UniversalDiffEq.jl, pure Julia) cannot be executed in
 * this image or on the GPU box (no Julia) and its own tests assert no numbers
 * (the test suite is maxiter=1 smoke tests), so this restatement is checked against closed forms,
 * complex-step derivatives of an independent NumPy twin (oracle/ude_oracle_np.py) and SciPy adaptive
 * solves instead of reference-produced golden vectors.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product path (universaldiffeq.jl_b200/csrc/libude_cuda.so) never does.
 *
 * What it restates (reference file:line, relative to /root/reference):
 *   loss assembly      src/LossFunctionConstructors.jl:17-98 (conditional likelihood), :175-192 / :500-540
 *                      (derivative matching), :196-232 / :545-601 (shooting), :261-305 / :639-708 (multiple
 *                      shooting); src/helpers.jl:16-76 and src/MultipleTimeSeries.jl:51-157 (default MSE
 *                      state-space loss) with src/LossFunctions.jl:17-28 weights
 *   segment semantics  src/ProcessModels.jl:72-141, :376-462; src/MultiProcessModel.jl:74-149, :204-280
 *                      (one ODE solve per interval / block, series index passed to the RHS)
 *   regulariser        src/Regularization.jl:19-36, :97-105
 *   covariates         src/covariates.jl:14-80 (piecewise linear, flat extrapolation)
 *   theta layout       src/helpers.jl:1-12, :491-556 (uhat d x N column-major first, then process_model)
 *   post-train states  src/LossFunctionConstructors.jl:237-258, :310-341, :604-631, :711-756
 * Third-party arithmetic restated from its published definition (source not in /root/reference):
 *   OrdinaryDiffEq Tsit5 tableau (compat ~6.93 / ^6.102.1, Project.toml:63) used at FIXED step (the north
 *   star's discretisation; the adaptive controller is not reproduced), classic RK4, Lux Dense/tanh MLP,
 *   and the discrete adjoint that ForwardDiffSensitivity+Zygote would return for that fixed-step map.
 */
#ifndef UDE_ORACLE_H
#define UDE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_RK4 = 0, ORC_TSIT5 = 1, ORC_EULER = 2 };
enum { ORC_LOSS_STATE_SPACE = 0, ORC_LOSS_MULTI_SHOOT = 1, ORC_LOSS_SHOOT = 2, ORC_LOSS_DERIV_MATCH = 3 };
enum { ORC_RHS_NODE = 0, ORC_RHS_NODE_TIME = 1, ORC_RHS_LV_UDE = 2, ORC_RHS_LV_UDE_ABS = 3,
       ORC_RHS_FISHERY = 4, ORC_RHS_NN_DECAY = 5, ORC_RHS_SERIES_GROWTH = 6 };

typedef struct orc_problem {
  /* model */
  int32_t d, n_cov, nn_in, nn_hidden, nn_out, rhs_kind, n_scalars, has_series_param;
  int32_t solver, substeps;
  int64_t off_w1, off_b1, off_w2, off_b2;   /* offsets inside process_model */
  int64_t off_scalar[8];
  int64_t off_series_param;
  int64_t n_pm;
  double time_scale, time_shift;
  /* loss */
  int32_t loss_kind, pred_length, proc_inv_dt, shoot_from_theta, shoot_first_scored, remove_ends;
  double preroll_eps;
  const double* A;                 /* d*d column-major: observation weight matrix */
  const double* B;                 /* d*d column-major: process weight matrix */
  /* data */
  int64_t n_series, n_cols, n_models, n_theta;
  const int64_t* starts;           /* 0-based first column of each series */
  const int64_t* lengths;
  const int64_t* model_of_series;
  const double* times;             /* n_cols */
  const double* data;              /* d x n_cols column-major */
  const int64_t* theta_base;       /* n_models */
  const int64_t* model_col0;       /* n_models+1 */
  const int64_t* model_series0;    /* n_models+1 */
  const double* obs_scale;         /* per model */
  const double* proc_scale;
  const double* shoot_weight;
  const double* reg_weight;
  const int64_t* knot_off;         /* n_series*n_cov+1 (NULL when n_cov==0) */
  const double* knot_t;
  const double* knot_x;
  const double* dm_u;              /* derivative matching: smoothed states d x n_cols */
  const double* dm_dudt;
  const uint8_t* skip_mask;        /* n_cols or NULL: interval ENDING at this column is skipped */
} orc_problem;

/* loss[n_models]; grad[n_theta] may be NULL (loss only). Returns 0, or <0 on bad input. */
int orc_loss_grad(const orc_problem* p, const double* theta, double* loss, double* grad, int n_threads);
/* predicted states at every column (see DESIGN.md "forward"), out is d x n_cols column-major */
int orc_forward(const orc_problem* p, const double* theta, double* out, int n_threads);
/* number of RK steps one loss evaluation takes (the unit of the throughput metric) */
int64_t orc_count_steps(const orc_problem* p);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
