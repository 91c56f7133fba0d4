"""Flat description of one batched loss-and-gradient problem (what crosses the C ABI).

This is the host-side mirror of ``ude_desc`` + the arrays passed to ``ude_set_data`` /
``ude_set_covariates`` / ``ude_set_targets`` (include/ude_cuda.h).  It restates the data layout of the
reference: ``theta = [uhat(:) | process_model | (empty groups)]`` with ``uhat`` d x N column-major
(/root/reference/src/helpers.jl:1-12), ``data`` d x N, ``times`` N, per-series ``starts`` / ``lengths``
(/root/reference/src/helpers.jl:491-556, made 0-based here), Lux Dense leaves ``weight`` (out x in,
column-major) then ``bias``.

Pure NumPy; no CUDA, no oracle imports.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

# solver
RK4, TSIT5, EULER = 0, 1, 2      # EULER: one RHS evaluation per interval (spline gradient matching)
SOLVER_STAGES = {RK4: 4, TSIT5: 6, EULER: 1}
# loss kinds
LOSS_STATE_SPACE, LOSS_MULTI_SHOOT, LOSS_SHOOT, LOSS_DERIV_MATCH = 0, 1, 2, 3
# right-hand-side catalog
RHS_NODE, RHS_NODE_TIME, RHS_LV_UDE, RHS_LV_UDE_ABS, RHS_FISHERY, RHS_NN_DECAY, RHS_SERIES_GROWTH = range(7)
RHS_NAMES = {RHS_NODE: "NODE", RHS_NODE_TIME: "NODE_TIME", RHS_LV_UDE: "LV_UDE", RHS_LV_UDE_ABS: "LV_UDE_ABS",
             RHS_FISHERY: "FISHERY", RHS_NN_DECAY: "NN_DECAY", RHS_SERIES_GROWTH: "SERIES_GROWTH"}

MAX_SCALARS = 8


def _f64(a, order="C"):
    return np.array(a, dtype=np.float64, order=order, copy=True)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


@dataclass
class Problem:
    # ---- model -------------------------------------------------------------------------------
    d: int
    nn_in: int
    nn_hidden: int
    nn_out: int
    rhs_kind: int = RHS_NODE
    n_cov: int = 0
    n_scalars: int = 0
    has_series_param: bool = False
    solver: int = RK4
    substeps: int = 4
    time_scale: float = 1.0
    time_shift: float = 0.0
    user_rhs: object = None          # rhs_kind >= 1000: the expression right-hand side (user_rhs.UserRhs) the kind stands for
    # process_model layout (offsets inside process_model); filled by ``default_layout``
    off_w1: int = 0
    off_b1: int = 0
    off_w2: int = 0
    off_b2: int = 0
    off_scalar: List[int] = field(default_factory=list)
    off_series_param: int = -1
    n_pm: int = 0
    # ---- loss --------------------------------------------------------------------------------
    loss_kind: int = LOSS_STATE_SPACE
    pred_length: int = 5
    proc_inv_dt: bool = False
    shoot_from_theta: bool = True
    shoot_first_scored: int = 2
    remove_ends: int = 0
    preroll_eps: float = 0.0
    A: Optional[np.ndarray] = None          # d x d observation weight (Sigma_obs^-1 or I)
    B: Optional[np.ndarray] = None          # d x d process weight
    # ---- data --------------------------------------------------------------------------------
    starts: Optional[np.ndarray] = None     # 0-based first column of each series
    lengths: Optional[np.ndarray] = None
    times: Optional[np.ndarray] = None      # n_cols
    data: Optional[np.ndarray] = None       # d x n_cols
    model_of_series: Optional[np.ndarray] = None
    obs_scale: Optional[np.ndarray] = None  # per model
    proc_scale: Optional[np.ndarray] = None
    shoot_weight: Optional[np.ndarray] = None
    reg_weight: Optional[np.ndarray] = None
    knot_off: Optional[np.ndarray] = None   # n_series*n_cov + 1
    knot_t: Optional[np.ndarray] = None
    knot_x: Optional[np.ndarray] = None
    dm_u: Optional[np.ndarray] = None       # derivative matching: smoothed states d x n_cols
    dm_dudt: Optional[np.ndarray] = None
    skip_mask: Optional[np.ndarray] = None  # uint8 n_cols: the interval ENDING at this column is skipped
    # derived by finalize()
    n_series: int = 0
    n_cols: int = 0
    n_models: int = 1
    n_theta: int = 0
    theta_base: Optional[np.ndarray] = None
    model_col0: Optional[np.ndarray] = None
    model_series0: Optional[np.ndarray] = None

    # ------------------------------------------------------------------------------------------
    def default_layout(self, n_series_param: int = 0) -> "Problem":
        """Lux order: layer_1.weight, layer_1.bias, layer_2.weight, layer_2.bias, then scalars, then the
        per-series vector (e.g. ``(NN = ..., r = zeros(m))``, docs/src/MultipleTimeSeries.md:57)."""
        H, din, dout = self.nn_hidden, self.nn_in, self.nn_out
        self.off_w1 = 0
        self.off_b1 = H * din
        self.off_w2 = self.off_b1 + H
        self.off_b2 = self.off_w2 + dout * H
        o = self.off_b2 + dout
        self.off_scalar = [o + k for k in range(self.n_scalars)]
        o += self.n_scalars
        if self.has_series_param:
            self.off_series_param = o
            o += n_series_param
        else:
            self.off_series_param = -1
        self.n_pm = o
        return self

    def finalize(self) -> "Problem":
        d = self.d
        self.starts = _i64(self.starts)
        self.lengths = _i64(self.lengths)
        self.n_series = int(self.starts.shape[0])
        self.times = _f64(self.times)
        self.n_cols = int(self.times.shape[0])
        self.data = np.asfortranarray(np.asarray(self.data, dtype=np.float64).reshape(d, self.n_cols, order="A"))
        if self.model_of_series is None:
            self.model_of_series = np.zeros(self.n_series, dtype=np.int64)
        self.model_of_series = _i64(self.model_of_series)
        M = int(self.model_of_series.max()) + 1 if self.n_series else 1
        self.n_models = M
        if np.any(np.diff(self.model_of_series) < 0):
            raise ValueError("series must be sorted by model")
        if np.any(self.starts[1:] != self.starts[:-1] + self.lengths[:-1]) or (self.n_series and self.starts[0] != 0):
            raise ValueError("series must tile the columns contiguously (sorted by series, then time)")
        ends = self.starts + self.lengths
        self.model_series0 = np.zeros(M + 1, dtype=np.int64)
        self.model_col0 = np.zeros(M + 1, dtype=np.int64)
        for m in range(M):
            idx = np.nonzero(self.model_of_series == m)[0]
            if idx.size == 0:
                raise ValueError(f"model {m} has no series")
            self.model_series0[m] = idx[0]
            self.model_col0[m] = self.starts[idx[0]]
        self.model_series0[M] = self.n_series
        self.model_col0[M] = int(ends[-1]) if self.n_series else 0
        if self.has_series_param and M > 1:
            n_sp = np.diff(self.model_series0)
            if np.any(n_sp != n_sp[0]):
                raise ValueError("per-series parameters need the same series count in every model")
        ncols_m = np.diff(self.model_col0)
        sizes = d * ncols_m + self.n_pm
        self.theta_base = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        self.n_theta = int(sizes.sum())
        one = np.ones(M)
        self.obs_scale = _f64(one if self.obs_scale is None else np.broadcast_to(self.obs_scale, (M,)))
        self.proc_scale = _f64(one if self.proc_scale is None else np.broadcast_to(self.proc_scale, (M,)))
        self.shoot_weight = _f64(one if self.shoot_weight is None else np.broadcast_to(self.shoot_weight, (M,)))
        self.reg_weight = _f64(0 * one if self.reg_weight is None else np.broadcast_to(self.reg_weight, (M,)))
        self.A = np.asfortranarray(np.eye(d) if self.A is None else np.asarray(self.A, dtype=np.float64).reshape(d, d))
        self.B = np.asfortranarray(np.eye(d) if self.B is None else np.asarray(self.B, dtype=np.float64).reshape(d, d))
        if self.n_cov > 0:
            self.knot_off = _i64(self.knot_off)
            self.knot_t = _f64(self.knot_t)
            self.knot_x = _f64(self.knot_x)
            if self.knot_off.shape[0] != self.n_series * self.n_cov + 1:
                raise ValueError("knot_off must have n_series*n_cov+1 entries")
        if self.loss_kind == LOSS_DERIV_MATCH:
            self.dm_u = np.asfortranarray(np.asarray(self.dm_u, dtype=np.float64).reshape(d, self.n_cols, order="A"))
            self.dm_dudt = np.asfortranarray(np.asarray(self.dm_dudt, dtype=np.float64).reshape(d, self.n_cols, order="A"))
        if self.skip_mask is not None:
            self.skip_mask = np.ascontiguousarray(np.asarray(self.skip_mask, dtype=np.uint8))
        if len(self.off_scalar) != self.n_scalars or self.n_scalars > MAX_SCALARS:
            raise ValueError("off_scalar / n_scalars mismatch")
        return self

    # ---- bookkeeping the benchmark metric needs ------------------------------------------------
    def pm_base(self, m: int = 0) -> int:
        return int(self.theta_base[m] + self.d * (self.model_col0[m + 1] - self.model_col0[m]))

    def uhat_view(self, theta: np.ndarray, m: int = 0) -> np.ndarray:
        n = int(self.model_col0[m + 1] - self.model_col0[m])
        b = int(self.theta_base[m])
        return theta[b:b + self.d * n].reshape(self.d, n, order="F")

    def pm_view(self, theta: np.ndarray, m: int = 0) -> np.ndarray:
        b = self.pm_base(m)
        return theta[b:b + self.n_pm]

    def count_segments(self) -> int:
        L = self.lengths
        if self.loss_kind == LOSS_STATE_SPACE:
            n = int(np.sum(np.maximum(L - 1, 0)))
            if self.skip_mask is not None:
                n -= int(sum(int(self.skip_mask[s + 1:s + l].sum()) for s, l in zip(self.starts, L)))
            return n
        if self.loss_kind == LOSS_MULTI_SHOOT:
            return int(np.sum((L + self.pred_length - 1) // self.pred_length))
        if self.loss_kind == LOSS_SHOOT:
            return int(self.n_series)
        return int(np.sum(np.maximum(L - 2 * self.remove_ends, 0)))

    def count_steps(self) -> int:
        """RK steps per loss evaluation: the unit of BASELINE.json's metric."""
        L = self.lengths
        S = self.substeps
        if self.loss_kind == LOSS_STATE_SPACE:
            n = int(np.sum(np.maximum(L - 1, 0)))
            if self.skip_mask is not None:
                n -= int(sum(int(self.skip_mask[s + 1:s + l].sum()) for s, l in zip(self.starts, L)))
            return n * S
        if self.loss_kind == LOSS_MULTI_SHOOT:
            pre = 1 if self.preroll_eps > 0 else 0
            nseg = (L + self.pred_length - 1) // self.pred_length
            return int(np.sum(np.maximum(L - 1, 0)) * S + np.sum(nseg) * pre)
        if self.loss_kind == LOSS_SHOOT:
            return int(np.sum(np.maximum(L - 1, 0)) * S)
        return 0

    def flops_per_step(self) -> float:
        """Algorithmic flops per RK trajectory-step, SURVEY.md 8(d): s*(3*F_mv + 5H + 2*dout) + comb,
        tanh counted as one flop, F_mv = 2H(din+dout), comb = 28d (RK4) / 84d (Tsit5)."""
        s = SOLVER_STAGES[self.solver]
        fmv = 2 * self.nn_hidden * (self.nn_in + self.nn_out)
        comb = {RK4: 28, TSIT5: 84, EULER: 4}[self.solver] * self.d
        return s * (3 * fmv + 5 * self.nn_hidden + 2 * self.nn_out) + comb

    def hbm_bytes_per_eval(self) -> float:
        """Algorithmic HBM bytes of one loss+grad evaluation: every column's data, times and
        uhat read once, the uhat gradient written once, covariate knots read once, theta/grad of the process
        model once."""
        b = 8.0 * (self.d * self.n_cols * 3 + self.n_cols)
        if self.n_cov:
            b += 16.0 * float(self.knot_t.shape[0])
        b += 16.0 * self.n_pm * self.n_models
        return b
