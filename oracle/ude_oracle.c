/* ude_oracle.c -- CPU oracle (scalar C, OpenMP over series) for the UDE loss+adjoint hot path.
 * TEST INFRASTRUCTURE ONLY; see ude_oracle.h for scope, citations and the "parity unpinned" notice.
 *
 * Structure (deliberately the naive one, nothing shared with the CUDA kernels):
 *   mlp_forward / mlp_vjp        Lux Chain(Dense(in,H,tanh), Dense(H,out))  src/ProcessModels.jl:379,427
 *   rhs_forward / rhs_vjp        the right-hand sides users of CustomDerivatives / NODE write
 *   covariates_at                src/covariates.jl:16-19, :45-48
 *   rk_step / rk_step_reverse    explicit RK stage loop and its discrete adjoint
 *   seg_state_space / seg_trajectory / col_deriv_match     one loss term each
 */
#include "ude_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 16
#define MAXS 6

/* ---- Butcher tableaux ------------------------------------------------------------------- */
typedef struct { int s; double c[MAXS]; double a[MAXS][MAXS]; double b[MAXS]; } tableau;

static const tableau TAB_RK4 = {
  4, {0.0, 0.5, 0.5, 1.0},
  {{0}, {0.5}, {0.0, 0.5}, {0.0, 0.0, 1.0}},
  {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0}};

/* Tsit5 (Tsitouras 2011) as in OrdinaryDiffEq's Tsit5ConstantCache; fixed-step use needs stages 1..6,
 * the update weights are row 7 of a (FSAL, b7 = 0).  SURVEY.md App. C.1. */
static const tableau TAB_TSIT5 = {
  6, {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0},
  {{0},
   {0.161},
   {-0.008480655492356989, 0.335480655492357},
   {2.8971530571054935, -6.359448489975075, 4.3622954328695815},
   {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},
   {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383}},
  {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};

/* explicit Euler: the spline-gradient-matching loss compares finite differences of the knot values with ONE right-hand-side
 * evaluation at the left knot (/root/reference/src/LossFunctionConstructors.jl:918-921) = one Euler step per interval */
static const tableau TAB_EULER = {1, {0.0}, {{0}}, {1.0}};

static const tableau* get_tab(int solver) { return solver == ORC_TSIT5 ? &TAB_TSIT5 : (solver == ORC_EULER ? &TAB_EULER : &TAB_RK4); }

/* ---- per-thread workspace ---------------------------------------------------------------- */
typedef struct {
  /* one record per RK stage of the current segment */
  double* U;   /* [n_stage_records][d]      stage input state */
  double* X;   /* [..][n_cov]               covariates at the stage time */
  double* x;   /* [..][nn_in]               MLP input */
  double* h;   /* [..][H]                   hidden activations */
  double* y;   /* [..][nn_out]              MLP output */
  double* tt;  /* [..]                      stage time */
  double* hs;  /* [n_steps]                 step sizes */
  double* seed;/* [(L+1)][d]                loss cotangent at each observation point */
  int64_t cap_stages, cap_steps, cap_pts;
  double* gpm;      /* [n_models? no: per-thread accumulators are handled by caller] */
} workspace;

static void ws_reserve(workspace* w, const orc_problem* p, int64_t n_steps, int64_t n_pts) {
  const tableau* T = get_tab(p->solver);
  int64_t n_st = n_steps * T->s;
  if (n_st > w->cap_stages) {
    free(w->U); free(w->X); free(w->x); free(w->h); free(w->y); free(w->tt);
    w->cap_stages = n_st + 16;
    w->U = (double*)malloc(sizeof(double) * w->cap_stages * p->d);
    w->X = (double*)malloc(sizeof(double) * w->cap_stages * (p->n_cov > 0 ? p->n_cov : 1));
    w->x = (double*)malloc(sizeof(double) * w->cap_stages * p->nn_in);
    w->h = (double*)malloc(sizeof(double) * w->cap_stages * p->nn_hidden);
    w->y = (double*)malloc(sizeof(double) * w->cap_stages * p->nn_out);
    w->tt = (double*)malloc(sizeof(double) * w->cap_stages);
  }
  if (n_steps > w->cap_steps) {
    free(w->hs); w->cap_steps = n_steps + 16;
    w->hs = (double*)malloc(sizeof(double) * w->cap_steps);
  }
  if (n_pts > w->cap_pts) {
    free(w->seed); w->cap_pts = n_pts + 16;
    w->seed = (double*)malloc(sizeof(double) * w->cap_pts * p->d);
  }
}
static void ws_free(workspace* w) {
  free(w->U); free(w->X); free(w->x); free(w->h); free(w->y); free(w->tt); free(w->hs); free(w->seed);
  memset(w, 0, sizeof(*w));
}

/* ---- covariates: src/covariates.jl:16 linear_interpolation(..., extrapolation_bc = Flat()) ---- */
static void covariates_at(const orc_problem* p, int64_t series, double t, double* X) {
  for (int v = 0; v < p->n_cov; ++v) {
    int64_t k0 = p->knot_off[series * p->n_cov + v], k1 = p->knot_off[series * p->n_cov + v + 1];
    const double* kt = p->knot_t + k0; const double* kx = p->knot_x + k0;
    int64_t n = k1 - k0;
    if (n <= 0) { X[v] = 0.0; continue; }
    if (t <= kt[0]) { X[v] = kx[0]; continue; }
    if (t >= kt[n - 1]) { X[v] = kx[n - 1]; continue; }
    int64_t lo = 0, hi = n - 1;            /* kt[lo] <= t < kt[hi] */
    while (hi - lo > 1) { int64_t mid = (lo + hi) / 2; if (kt[mid] <= t) lo = mid; else hi = mid; }
    double f = (t - kt[lo]) / (kt[lo + 1] - kt[lo]);
    X[v] = (1.0 - f) * kx[lo] + f * kx[lo + 1];
  }
}

/* ---- MLP: y = W2 tanh(W1 x + b1) + b2; weights column-major (out x in), SURVEY App. B ---- */
static void mlp_forward(const orc_problem* p, const double* pm, const double* x, double* h, double* y) {
  const double* W1 = pm + p->off_w1; const double* b1 = pm + p->off_b1;
  const double* W2 = pm + p->off_w2; const double* b2 = pm + p->off_b2;
  int H = p->nn_hidden, din = p->nn_in, dout = p->nn_out;
  for (int i = 0; i < H; ++i) {
    double z = b1[i];
    for (int k = 0; k < din; ++k) z += W1[i + (int64_t)H * k] * x[k];
    h[i] = tanh(z);
  }
  for (int j = 0; j < dout; ++j) {
    double s = b2[j];
    for (int i = 0; i < H; ++i) s += W2[j + (int64_t)dout * i] * h[i];
    y[j] = s;
  }
}
/* given ybar: accumulate parameter cotangents into gpm (may be NULL) and return xbar */
static void mlp_vjp(const orc_problem* p, const double* pm, const double* x, const double* h,
                    const double* ybar, double* xbar, double* gpm) {
  const double* W1 = pm + p->off_w1; const double* W2 = pm + p->off_w2;
  int H = p->nn_hidden, din = p->nn_in, dout = p->nn_out;
  for (int k = 0; k < din; ++k) xbar[k] = 0.0;
  for (int i = 0; i < H; ++i) {
    double hb = 0.0;
    for (int j = 0; j < dout; ++j) hb += W2[j + (int64_t)dout * i] * ybar[j];
    double zb = hb * (1.0 - h[i] * h[i]);
    if (gpm) {
      for (int j = 0; j < dout; ++j) gpm[p->off_w2 + j + (int64_t)dout * i] += ybar[j] * h[i];
      gpm[p->off_b1 + i] += zb;
      for (int k = 0; k < din; ++k) gpm[p->off_w1 + i + (int64_t)H * k] += zb * x[k];
    }
    for (int k = 0; k < din; ++k) xbar[k] += W1[i + (int64_t)H * k] * zb;
  }
  if (gpm) for (int j = 0; j < dout; ++j) gpm[p->off_b2 + j] += ybar[j];
}

static double sgn(double v) { return v < 0.0 ? -1.0 : 1.0; }

/* ---- right-hand sides ---------------------------------------------------- */
static void rhs_forward(const orc_problem* p, const double* pm, int64_t series_local, const double* u,
                        const double* X, double t, double* du, double* x, double* h, double* y) {
  int d = p->d;
  for (int j = 0; j < d; ++j) x[j] = u[j];
  switch (p->rhs_kind) {
    case ORC_RHS_NODE:      /* src/ProcessModels.jl:387-390, :434-437; src/MultiProcessModel.jl:215-218 */
    case ORC_RHS_FISHERY:   /* docs/src/examples.md:156-169: NN(vcat(u,X)) */
      for (int v = 0; v < p->n_cov; ++v) x[d + v] = X[v];
      break;
    case ORC_RHS_NODE_TIME: /* docs/src/examples.md:44-49: NN([u..., X..., t/50 - 1]) */
      for (int v = 0; v < p->n_cov; ++v) x[d + v] = X[v];
      x[d + p->n_cov] = p->time_scale * t + p->time_shift;
      break;
    default: break;          /* LV_UDE, LV_UDE_ABS, NN_DECAY, SERIES_GROWTH: NN(u) */
  }
  mlp_forward(p, pm, x, h, y);
  const double* sc[8];
  for (int k = 0; k < p->n_scalars && k < 8; ++k) sc[k] = pm + p->off_scalar[k];
  switch (p->rhs_kind) {
    case ORC_RHS_NODE:
      for (int j = 0; j < d; ++j) du[j] = y[j];
      break;
    case ORC_RHS_NODE_TIME:
      for (int j = 0; j < d; ++j) du[j] = (j < p->nn_out) ? y[j] : -(*sc[j - p->nn_out]) * u[j];
      break;
    case ORC_RHS_LV_UDE: {  /* test/UDETests.jl:20-28 (r, m, theta) and :82-86 (+ beta[1:2] * X[1]) */
      double r = *sc[0], m = *sc[1], th = *sc[2], C = y[0];
      du[0] = r * u[0] - C;
      du[1] = th * C - m * u[1];
      if (p->n_scalars >= 5) { du[0] += (*sc[3]) * X[0]; du[1] += (*sc[4]) * X[0]; }
    } break;
    case ORC_RHS_LV_UDE_ABS: { /* README.md:52-62: abs on NN output and on (r, k, theta, m); p order r,k,m,theta */
      double r = fabs(*sc[0]), k = fabs(*sc[1]), m = fabs(*sc[2]), th = fabs(*sc[3]), C = fabs(y[0]);
      du[0] = r * u[0] * (1.0 - u[0] / k) - C;
      du[1] = th * C - m * u[1];
    } break;
    case ORC_RHS_FISHERY: {  /* docs/src/examples.md:163-166, p order q, r, K */
      double q = *sc[0], r = *sc[1], K = *sc[2];
      du[0] = y[0];
      du[1] = r * u[1] * (1.0 - u[1] / K) - q * u[0];
    } break;
    case ORC_RHS_NN_DECAY: { /* test/MultiUDE.jl:17-19: NN(u) .- m*u */
      double m = *sc[0];
      for (int j = 0; j < d; ++j) du[j] = y[j] - m * u[j];
    } break;
    case ORC_RHS_SERIES_GROWTH: { /* docs/src/MultipleTimeSeries.md:51-55: r[i] .* u .* NN(u)[1] */
      double r = pm[p->off_series_param + series_local];
      for (int j = 0; j < d; ++j) du[j] = r * u[j] * y[0];
    } break;
  }
}

/* ubar += (df/du)^T kbar ; gpm += (df/dtheta)^T kbar */
static void rhs_vjp(const orc_problem* p, const double* pm, int64_t series_local, const double* u,
                    const double* X, const double* x, const double* h, const double* y,
                    const double* kbar, double* ubar, double* gpm) {
  int d = p->d;
  double ybar[MAXD]; double xbar[2 * MAXD + 2];
  for (int j = 0; j < p->nn_out; ++j) ybar[j] = 0.0;
  const double* sc[8]; double* gs[8];
  for (int k = 0; k < p->n_scalars && k < 8; ++k) { sc[k] = pm + p->off_scalar[k]; gs[k] = gpm ? gpm + p->off_scalar[k] : NULL; }
  switch (p->rhs_kind) {
    case ORC_RHS_NODE:
      for (int j = 0; j < d; ++j) ybar[j] = kbar[j];
      break;
    case ORC_RHS_NODE_TIME:
      for (int j = 0; j < d; ++j) {
        if (j < p->nn_out) ybar[j] = kbar[j];
        else { ubar[j] += -(*sc[j - p->nn_out]) * kbar[j]; if (gpm) *gs[j - p->nn_out] += -u[j] * kbar[j]; }
      }
      break;
    case ORC_RHS_LV_UDE: {
      double r = *sc[0], m = *sc[1], th = *sc[2], C = y[0];
      ybar[0] = -kbar[0] + th * kbar[1];
      ubar[0] += r * kbar[0];
      ubar[1] += -m * kbar[1];
      if (gpm) {
        *gs[0] += u[0] * kbar[0]; *gs[1] += -u[1] * kbar[1]; *gs[2] += C * kbar[1];
        if (p->n_scalars >= 5) { *gs[3] += X[0] * kbar[0]; *gs[4] += X[0] * kbar[1]; }
      }
    } break;
    case ORC_RHS_LV_UDE_ABS: {
      double r = fabs(*sc[0]), k = fabs(*sc[1]), m = fabs(*sc[2]), th = fabs(*sc[3]), C = fabs(y[0]);
      ybar[0] = sgn(y[0]) * (-kbar[0] + th * kbar[1]);
      ubar[0] += r * (1.0 - 2.0 * u[0] / k) * kbar[0];
      ubar[1] += -m * kbar[1];
      if (gpm) {
        *gs[0] += sgn(*sc[0]) * u[0] * (1.0 - u[0] / k) * kbar[0];
        *gs[1] += sgn(*sc[1]) * r * u[0] * u[0] / (k * k) * kbar[0];
        *gs[2] += sgn(*sc[2]) * (-u[1]) * kbar[1];
        *gs[3] += sgn(*sc[3]) * C * kbar[1];
      }
    } break;
    case ORC_RHS_FISHERY: {
      double q = *sc[0], r = *sc[1], K = *sc[2];
      ybar[0] = kbar[0];
      ubar[0] += -q * kbar[1];
      ubar[1] += r * (1.0 - 2.0 * u[1] / K) * kbar[1];
      if (gpm) {
        *gs[0] += -u[0] * kbar[1];
        *gs[1] += u[1] * (1.0 - u[1] / K) * kbar[1];
        *gs[2] += r * u[1] * u[1] / (K * K) * kbar[1];
      }
    } break;
    case ORC_RHS_NN_DECAY: {
      double m = *sc[0]; double acc = 0.0;
      for (int j = 0; j < d; ++j) { ybar[j] = kbar[j]; ubar[j] += -m * kbar[j]; acc += u[j] * kbar[j]; }
      if (gpm) *gs[0] += -acc;
    } break;
    case ORC_RHS_SERIES_GROWTH: {
      double r = pm[p->off_series_param + series_local]; double acc = 0.0;
      for (int j = 0; j < d; ++j) { acc += u[j] * kbar[j]; ubar[j] += r * y[0] * kbar[j]; }
      ybar[0] = r * acc;
      if (gpm) gpm[p->off_series_param + series_local] += y[0] * acc;
    } break;
  }
  mlp_vjp(p, pm, x, h, ybar, xbar, gpm);
  for (int j = 0; j < d; ++j) ubar[j] += xbar[j];   /* the first d MLP inputs are always u */
}

/* ---- one explicit RK step, recording what the reverse sweep needs ------------------------- */
static void rk_step(const orc_problem* p, const double* pm, int64_t series_global, int64_t series_local,
                    double* u, double t, double hstep, workspace* w, int64_t rec0 /* first stage record or -1 */) {
  const tableau* T = get_tab(p->solver);
  int d = p->d;
  double k[MAXS][MAXD], U[MAXD], X[MAXD], xb[2 * MAXD + 2], hb[4096], yb[MAXD];
  for (int i = 0; i < T->s; ++i) {
    for (int j = 0; j < d; ++j) {
      double acc = 0.0;
      for (int l = 0; l < i; ++l) acc += T->a[i][l] * k[l][j];
      U[j] = u[j] + hstep * acc;
    }
    double ts = t + T->c[i] * hstep;
    if (p->n_cov > 0) covariates_at(p, series_global, ts, X);
    double *x = xb, *hh = hb, *y = yb;
    if (rec0 >= 0) {
      int64_t r = rec0 + i;
      memcpy(w->U + r * d, U, sizeof(double) * d);
      if (p->n_cov > 0) memcpy(w->X + r * p->n_cov, X, sizeof(double) * p->n_cov);
      x = w->x + r * p->nn_in; hh = w->h + r * p->nn_hidden; y = w->y + r * p->nn_out; w->tt[r] = ts;
    }
    rhs_forward(p, pm, series_local, U, X, ts, k[i], x, hh, y);
  }
  for (int j = 0; j < d; ++j) {
    double acc = 0.0;
    for (int i = 0; i < T->s; ++i) acc += T->b[i] * k[i][j];
    u[j] += hstep * acc;
  }
}

/* kbar_i = h b_i lam + h sum_{j>i} a_ji Ubar_j ; (Ubar_i, thetabar) = VJP_f ; lam += sum Ubar_i */
static void rk_step_reverse(const orc_problem* p, const double* pm, int64_t series_local, double* lam,
                            double hstep, const workspace* w, int64_t rec0, double* gpm) {
  const tableau* T = get_tab(p->solver);
  int d = p->d;
  double Ub[MAXS][MAXD];
  for (int i = T->s - 1; i >= 0; --i) {
    double kb[MAXD];
    for (int j = 0; j < d; ++j) {
      double acc = T->b[i] * lam[j];
      for (int l = i + 1; l < T->s; ++l) acc += T->a[l][i] * Ub[l][j];
      kb[j] = hstep * acc;
      Ub[i][j] = 0.0;
    }
    int64_t r = rec0 + i;
    rhs_vjp(p, pm, series_local, w->U + r * d, w->X + r * (p->n_cov > 0 ? p->n_cov : 0), w->x + r * p->nn_in,
            w->h + r * p->nn_hidden, w->y + r * p->nn_out, kb, Ub[i], gpm);
  }
  for (int j = 0; j < d; ++j) { double acc = 0.0; for (int i = 0; i < T->s; ++i) acc += Ub[i][j]; lam[j] += acc; }
}

/* quadratic form r' M r and (M + M') r */
static double quad(const double* M, const double* r, int d, double* out_sym) {
  double q = 0.0;
  for (int a = 0; a < d; ++a) {
    double s = 0.0, ssym = 0.0;
    for (int b = 0; b < d; ++b) { s += M[a + d * b] * r[b]; ssym += (M[a + d * b] + M[b + d * a]) * r[b]; }
    q += r[a] * s; out_sym[a] = ssym;
  }
  return q;
}

/* ---- loss terms ---------------------------------------------------------------------------- */
typedef struct { const orc_problem* p; const double* theta; double* grad; int64_t model; int64_t series;
                 int64_t series_local; const double* uhat; double* guhat; const double* pm; double* gpm; } ctx;

/* observation term of one column: (y-uhat)' A (y-uhat) * obs_scale   (LFC.jl:30-34; helpers.jl:23-27) */
static double obs_term(const ctx* c, int64_t col_local, int64_t col_global) {
  const orc_problem* p = c->p; int d = p->d;
  double e[MAXD], sym[MAXD];
  for (int j = 0; j < d; ++j) e[j] = p->data[j + d * col_global] - c->uhat[j + d * col_local];
  double sc = p->obs_scale[c->model];
  double q = quad(p->A, e, d, sym);
  if (c->guhat) for (int j = 0; j < d; ++j) c->guhat[j + d * col_local] += -sc * sym[j];
  return sc * q;
}

/* process term of the interval ending at column col (LFC.jl:38-44; helpers.jl:31-37) */
static double seg_state_space(const ctx* c, int64_t col_local, int64_t col_global, workspace* w) {
  const orc_problem* p = c->p; int d = p->d; const tableau* T = get_tab(p->solver);
  int S = p->substeps;
  ws_reserve(w, p, S, 2);
  double u[MAXD];
  for (int j = 0; j < d; ++j) u[j] = c->uhat[j + d * (col_local - 1)];
  double t0 = p->times[col_global - 1], dt = p->times[col_global] - t0, hs = dt / S;
  for (int n = 0; n < S; ++n) rk_step(p, c->pm, c->series, c->series_local, u, t0 + n * hs, hs, w, c->grad ? (int64_t)n * T->s : -1);
  double r[MAXD], sym[MAXD];
  for (int j = 0; j < d; ++j) r[j] = c->uhat[j + d * col_local] - u[j];
  double sc = p->proc_scale[c->model] * (p->proc_inv_dt ? 1.0 / dt : 1.0);
  double q = quad(p->B, r, d, sym);
  if (c->grad) {
    double lam[MAXD];
    for (int j = 0; j < d; ++j) { lam[j] = -sc * sym[j]; c->guhat[j + d * col_local] += sc * sym[j]; }
    for (int n = S - 1; n >= 0; --n) rk_step_reverse(p, c->pm, c->series_local, lam, hs, w, (int64_t)n * T->s, c->gpm);
    for (int j = 0; j < d; ++j) c->guhat[j + d * (col_local - 1)] += lam[j];
  }
  return sc * q;
}

/* shooting / multiple-shooting block starting at column c0 with L intervals (LFC.jl:261-305, :639-687, :196-232, :545-601).
 * traj_out (d x n_cols global, may be NULL): predicted state at every point of the block. */
static double seg_trajectory(const ctx* c, int64_t c0_local, int64_t c0_global, int64_t L, int from_theta,
                             double eps, int64_t first_scored, double wpt, workspace* w, double* traj_out) {
  const orc_problem* p = c->p; int d = p->d; const tableau* T = get_tab(p->solver);
  int S = p->substeps; int pre = eps > 0.0 ? 1 : 0;
  int64_t n_steps = pre + L * S;
  ws_reserve(w, p, n_steps, L + 1);
  double u[MAXD]; double loss = 0.0;
  for (int j = 0; j < d; ++j) u[j] = from_theta ? c->uhat[j + d * c0_local] : p->data[j + d * c0_global];
  int64_t step = 0;
  int rec = c->grad != NULL;
  if (pre) {  /* LFC.jl:268: tspan = (tsteps[1] - 10^-6, tsteps[end]); restated as ONE RK step of length eps */
    w->hs[step] = eps;
    rk_step(p, c->pm, c->series, c->series_local, u, p->times[c0_global] - eps, eps, w, rec ? step * T->s : -1);
    ++step;
  }
  for (int64_t i = 0; i <= L; ++i) {
    if (traj_out) for (int j = 0; j < d; ++j) traj_out[j + d * (c0_global + i)] = u[j];
    for (int j = 0; j < d; ++j) w->seed[i * d + j] = 0.0;
    if (i >= first_scored) {
      for (int j = 0; j < d; ++j) {
        double r = p->data[j + d * (c0_global + i)] - u[j];
        loss += wpt * r * r;
        w->seed[i * d + j] = -2.0 * wpt * r;
      }
    }
    if (i == L) break;
    double t0 = p->times[c0_global + i], hs = (p->times[c0_global + i + 1] - t0) / S;
    for (int n = 0; n < S; ++n) {
      w->hs[step] = hs;
      rk_step(p, c->pm, c->series, c->series_local, u, t0 + n * hs, hs, w, rec ? step * T->s : -1);
      ++step;
    }
  }
  if (c->grad) {
    double lam[MAXD];
    for (int j = 0; j < d; ++j) lam[j] = 0.0;
    for (int64_t i = L; i >= 0; --i) {
      for (int j = 0; j < d; ++j) lam[j] += w->seed[i * d + j];
      if (i == 0) break;
      for (int n = S - 1; n >= 0; --n) { --step; rk_step_reverse(p, c->pm, c->series_local, lam, w->hs[step], w, step * T->s, c->gpm); }
    }
    if (pre) { --step; rk_step_reverse(p, c->pm, c->series_local, lam, w->hs[step], w, step * T->s, c->gpm); }
    if (from_theta) for (int j = 0; j < d; ++j) c->guhat[j + d * c0_local] += lam[j];
  }
  return loss;
}

/* derivative matching term of one column (LFC.jl:181-186, :525-533) */
static double col_deriv_match(const ctx* c, int64_t col_global) {
  const orc_problem* p = c->p; int d = p->d;
  double X[MAXD], x[2 * MAXD + 2], hh[4096], y[MAXD], f[MAXD], kb[MAXD], ub[MAXD];
  double t = p->times[col_global];
  if (p->n_cov > 0) covariates_at(p, c->series, t, X);
  const double* u = p->dm_u + d * col_global;
  rhs_forward(p, c->pm, c->series_local, u, X, t, f, x, hh, y);
  double loss = 0.0;
  for (int j = 0; j < d; ++j) { double r = p->dm_dudt[j + d * col_global] - f[j]; loss += r * r; kb[j] = -2.0 * r; ub[j] = 0.0; }
  if (c->grad) rhs_vjp(p, c->pm, c->series_local, u, X, x, hh, y, kb, ub, c->gpm);
  return loss;
}

static double series_loss(const orc_problem* p, const double* theta, double* grad, int64_t s, workspace* w,
                          double* gpm_thread /* n_models? indexed by caller */, double* traj_out) {
  int d = p->d;
  int64_t m = p->model_of_series ? p->model_of_series[s] : 0;
  int64_t c0 = p->starts[s], len = p->lengths[s];
  int64_t ncols_m = p->model_col0[m + 1] - p->model_col0[m];
  ctx c; c.p = p; c.theta = theta; c.grad = grad; c.model = m; c.series = s; c.series_local = s - p->model_series0[m];
  c.uhat = theta + p->theta_base[m] - (int64_t)d * p->model_col0[m];          /* index with GLOBAL column */
  c.guhat = grad ? grad + p->theta_base[m] - (int64_t)d * p->model_col0[m] : NULL;
  c.pm = theta + p->theta_base[m] + (int64_t)d * ncols_m;
  c.gpm = gpm_thread;
  double L = 0.0;
  switch (p->loss_kind) {
    case ORC_LOSS_STATE_SPACE:
      for (int64_t t = 0; t < len; ++t) {
        L += obs_term(&c, c0 + t, c0 + t);
        if (traj_out && t == 0) for (int j = 0; j < d; ++j) traj_out[j + d * c0] = c.uhat[j + d * c0];
        if (t >= 1 && !(p->skip_mask && p->skip_mask[c0 + t])) {
          if (traj_out) {   /* one-step-ahead prediction (src/ModelTesting.jl:181-195) */
            double u[MAXD]; for (int j = 0; j < d; ++j) u[j] = c.uhat[j + d * (c0 + t - 1)];
            double t0 = p->times[c0 + t - 1], hs = (p->times[c0 + t] - t0) / p->substeps;
            for (int n = 0; n < p->substeps; ++n) rk_step(p, c.pm, s, c.series_local, u, t0 + n * hs, hs, w, -1);
            for (int j = 0; j < d; ++j) traj_out[j + d * (c0 + t)] = u[j];
          } else L += seg_state_space(&c, c0 + t, c0 + t, w);
        }
      }
      break;
    case ORC_LOSS_MULTI_SHOOT: {
      double wpt = 1.0 / ((double)d * (double)len);            /* LFC.jl:295, :679: / length(data) */
      for (int64_t t = 0; t < len; t += p->pred_length) {       /* LFC.jl:281-287, :665-673 */
        int64_t Lb = (len - 1 - t >= p->pred_length) ? p->pred_length : (len - 1 - t);
        L += seg_trajectory(&c, c0 + t, c0 + t, Lb, 1, p->preroll_eps, 0, wpt, w, traj_out);
      }
    } break;
    case ORC_LOSS_SHOOT:
      L += seg_trajectory(&c, c0, c0, len - 1, p->shoot_from_theta, 0.0, p->shoot_first_scored,
                          p->shoot_weight[m], w, traj_out);
      break;
    case ORC_LOSS_DERIV_MATCH:
      for (int64_t t = p->remove_ends; t < len - p->remove_ends; ++t) L += col_deriv_match(&c, c0 + t);
      break;
  }
  return L;
}

static int check(const orc_problem* p) {
  if (p->d < 1 || p->d > MAXD || p->nn_out > MAXD || p->nn_in > 2 * MAXD + 2 || p->nn_hidden > 4096) return -1;
  if (p->n_cov > MAXD || p->substeps < 1 || p->n_models < 1) return -1;
  return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int orc_loss_grad(const orc_problem* p, const double* theta, double* loss, double* grad, int n_threads) {
  if (check(p)) return -1;
  if (n_threads < 1) n_threads = 1;
  int64_t M = p->n_models, P = p->n_pm;
  for (int64_t m = 0; m < M; ++m) loss[m] = 0.0;
  if (grad) memset(grad, 0, sizeof(double) * p->n_theta);
  /* per-thread accumulators, combined in thread order => run-to-run deterministic for a fixed n_threads */
  double* tl = (double*)calloc((size_t)n_threads * M, sizeof(double));
  double* tg = grad ? (double*)calloc((size_t)n_threads * M * P, sizeof(double)) : NULL;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    workspace w; memset(&w, 0, sizeof(w));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t s = 0; s < p->n_series; ++s) {
      int64_t m = p->model_of_series ? p->model_of_series[s] : 0;
      /* uhat gradient entries are private to a series except nothing: columns belong to exactly one series */
      tl[(size_t)tid * M + m] += series_loss(p, theta, grad, s, &w, tg ? tg + ((size_t)tid * M + m) * P : NULL, NULL);
    }
    ws_free(&w);
  }
  for (int64_t m = 0; m < M; ++m) {
    int64_t ncols_m = p->model_col0[m + 1] - p->model_col0[m];
    const double* pm = theta + p->theta_base[m] + (int64_t)p->d * ncols_m;
    double* gpm = grad ? grad + p->theta_base[m] + (int64_t)p->d * ncols_m : NULL;
    for (int t = 0; t < n_threads; ++t) {
      loss[m] += tl[(size_t)t * M + m];
      if (grad) for (int64_t k = 0; k < P; ++k) gpm[k] += tg[((size_t)t * M + m) * P + k];
    }
    /* L2 on the two weight matrices only (src/Regularization.jl:19-36, :97-105); biases excluded */
    double rw = p->reg_weight ? p->reg_weight[m] : 0.0;
    if (rw != 0.0) {
      double acc = 0.0;
      int64_t n1 = (int64_t)p->nn_hidden * p->nn_in, n2 = (int64_t)p->nn_out * p->nn_hidden;
      for (int64_t k = 0; k < n1; ++k) { double v = pm[p->off_w1 + k]; acc += v * v; if (grad) gpm[p->off_w1 + k] += 2.0 * rw * v; }
      for (int64_t k = 0; k < n2; ++k) { double v = pm[p->off_w2 + k]; acc += v * v; if (grad) gpm[p->off_w2 + k] += 2.0 * rw * v; }
      loss[m] += rw * acc;
    }
  }
  free(tl); free(tg);
  return 0;
}

int orc_forward(const orc_problem* p, const double* theta, double* out, int n_threads) {
  if (check(p)) return -1;
  if (n_threads < 1) n_threads = 1;
  memset(out, 0, sizeof(double) * p->d * p->n_cols);
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
  {
    workspace w; memset(&w, 0, sizeof(w));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t s = 0; s < p->n_series; ++s) (void)series_loss(p, theta, NULL, s, &w, NULL, out);
    ws_free(&w);
  }
  return 0;
}

int64_t orc_count_steps(const orc_problem* p) {
  int64_t n = 0; int pre = p->preroll_eps > 0.0 ? 1 : 0;
  for (int64_t s = 0; s < p->n_series; ++s) {
    int64_t len = p->lengths[s], c0 = p->starts[s];
    switch (p->loss_kind) {
      case ORC_LOSS_STATE_SPACE:
        for (int64_t t = 1; t < len; ++t) if (!(p->skip_mask && p->skip_mask[c0 + t])) n += p->substeps;
        break;
      case ORC_LOSS_MULTI_SHOOT:
        for (int64_t t = 0; t < len; t += p->pred_length) {
          int64_t Lb = (len - 1 - t >= p->pred_length) ? p->pred_length : (len - 1 - t);
          n += pre + Lb * p->substeps;
        }
        break;
      case ORC_LOSS_SHOOT: n += (len - 1) * p->substeps; break;
      default: break;
    }
  }
  return n;
}
