"""Host-side mirror of the reference's model constructors for the hot path.

Same names, argument meaning and defaults as the Julia API; pandas DataFrames stand in for DataFrames.jl:
  SimpleNeuralNetwork      /root/reference/src/NeuralNetworkConstructors.jl:20-30
  CustomDerivatives        /root/reference/src/Models.jl:115-155 (+ covariates :248-288)
  NODE                     /root/reference/src/Models.jl:746-840
  MultiNODE                /root/reference/src/MultipleTimeSeries.jl:181-279
  MultiCustomDerivatives   /root/reference/src/MultipleTimeSeries.jl:301-362
  process_data / process_multi_data / interpolate_covariates   src/helpers.jl:491-556, src/covariates.jl:2-80
A Julia closure `derivs!` cannot run inside a hand-written kernel, so `derivs` is an entry of the RHS catalog
(rhs.py) instead of a function.  `ode_solver` takes the markers CudaRK4(substeps) /
CudaTsit5(substeps) (fixed-step replacements of OrdinaryDiffEq's adaptive Tsit5()).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np
import pandas as pd

from . import problem as P
from .synthetic import glorot_f32


# ---- solver markers ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class CudaRK4:
    substeps: int = 4
    kind: int = P.RK4


@dataclass(frozen=True)
class CudaTsit5:
    substeps: int = 4
    kind: int = P.TSIT5


# ---- RHS catalog (what `derivs!` may be) ---------------------------------------------------------------------------
@dataclass(frozen=True)
class Rhs:
    kind: int
    nn_in: int
    nn_out: int
    scalar_names: tuple = ()
    series_param: Optional[str] = None
    time_scale: float = 1.0
    time_shift: float = 0.0
    name: str = ""


def node_rhs(d, n_cov=0):
    return Rhs(P.RHS_NODE, d + n_cov, d, name="NODE")


def node_time_rhs(d, n_cov, nn_out, time_scale=1.0 / 50.0, time_shift=-1.0, decay_names=()):
    """du[1:nn_out] = NN([u; X; a t + b]); remaining states decay linearly (docs/src/examples.md:44-49)."""
    return Rhs(P.RHS_NODE_TIME, d + n_cov + 1, nn_out, tuple(decay_names), None, time_scale, time_shift, "NODE_TIME")


LV_UDE = Rhs(P.RHS_LV_UDE, 2, 1, ("r", "m", "theta"), name="LV_UDE")                       # test/UDETests.jl:20-28
LV_UDE_X = Rhs(P.RHS_LV_UDE, 2, 1, ("r", "m", "theta", "beta1", "beta2"), name="LV_UDE_X")   # test/UDETests.jl:82-86
LV_UDE_ABS = Rhs(P.RHS_LV_UDE_ABS, 2, 1, ("r", "k", "m", "theta"), name="LV_UDE_ABS")      # README.md:52-62
FISHERY = Rhs(P.RHS_FISHERY, 3, 1, ("q", "r", "K"), name="FISHERY")                         # docs/src/examples.md:163-166


def nn_decay_rhs(d=1):
    return Rhs(P.RHS_NN_DECAY, d, d, ("m",), name="NN_DECAY")                               # test/MultiUDE.jl:17-19


def series_growth_rhs(d=1):
    return Rhs(P.RHS_SERIES_GROWTH, d, 1, (), "r", name="SERIES_GROWTH")                    # docs/src/MultipleTimeSeries.md:51-55


# ---- neural network ---------------------------------------------------------------------------------------------------
@dataclass
class SimpleNN:
    inputs: int
    outputs: int
    hidden: int

    def __call__(self, x, parameters):
        W1, b1, W2, b2 = (parameters[k] for k in ("layer_1.weight", "layer_1.bias", "layer_2.weight", "layer_2.bias"))
        return W2 @ np.tanh(W1 @ np.asarray(x, dtype=np.float64) + b1) + b2


def SimpleNeuralNetwork(inputs, outputs, hidden=10, nonlinearity="tanh", seed=123):
    """One-hidden-layer tanh MLP, Lux-style init (glorot-uniform Float32 weights, zero biases)."""
    if nonlinearity != "tanh":
        raise ValueError("the CUDA hot path implements tanh only (the activation of every shipped constructor)")
    rng = np.random.default_rng(seed)
    params = {"layer_1.weight": glorot_f32(rng, hidden, inputs), "layer_1.bias": np.zeros(hidden),
              "layer_2.weight": glorot_f32(rng, outputs, hidden), "layer_2.bias": np.zeros(outputs)}
    return SimpleNN(inputs, outputs, hidden), params


# ---- data handling -------------------------------------------------------------------------------------------------
def process_data(data: pd.DataFrame, time_column_name):
    df = data.sort_values(time_column_name, kind="stable").reset_index(drop=True)
    times = df[time_column_name].to_numpy(dtype=np.float64)
    varnames = [c for c in df.columns if c != time_column_name]
    mat = np.asfortranarray(df[varnames].to_numpy(dtype=np.float64).T)
    N, dims = len(times), len(varnames)
    T = float(times[-1] - times[0])
    return N, dims, T, times, mat, df, varnames


def process_multi_data(data: pd.DataFrame, time_column_name, series_column_name):
    df = data.copy()
    labels = sorted(pd.unique(df[series_column_name]))
    code = {lab: i for i, lab in enumerate(labels)}
    df["_series_code"] = df[series_column_name].map(code)
    df = df.sort_values(["_series_code", time_column_name], kind="stable").reset_index(drop=True)
    times = df[time_column_name].to_numpy(dtype=np.float64)
    varnames = [c for c in df.columns if c not in (time_column_name, series_column_name, "_series_code")]
    mat = np.asfortranarray(df[varnames].to_numpy(dtype=np.float64).T)
    codes = df["_series_code"].to_numpy()
    starts = np.searchsorted(codes, np.arange(len(labels)), side="left").astype(np.int64)
    lengths = np.diff(np.append(starts, len(codes))).astype(np.int64)
    T = float(times.max())                      # src/helpers.jl:545: T = times[argmax(times)]
    return len(times), T, len(varnames), mat, times, df.drop(columns="_series_code"), starts, lengths, varnames, labels


def interpolate_covariates(X: pd.DataFrame, time_column_name, series_column_name=None, labels=None,
                           variable_column_name=None, value_column_name=None):
    """-> (n_cov, knot_off, knot_t, knot_x): CSR knots per (series, variable); wide or long format (src/covariates.jl).
    One sort of the whole frame (the reference filters the frame once per series, O(series x rows), src/covariates.jl:30-44)."""
    n_rows = len(X)
    if series_column_name is None:
        s_code, n_series = np.zeros(n_rows, dtype=np.int64), 1
    else:
        code = {lab: i for i, lab in enumerate(labels)}
        s_code = X[series_column_name].map(code).to_numpy(dtype=np.float64)
        n_series = len(labels)
    t = X[time_column_name].to_numpy(dtype=np.float64)
    if variable_column_name is None:          # wide: one column per covariate
        cols = [c for c in X.columns if c not in (time_column_name, series_column_name)]
        n_cov = len(cols)
        s_long = np.tile(s_code, n_cov)
        v_long = np.repeat(np.arange(n_cov), n_rows)
        t_long = np.tile(t, n_cov)
        x_long = np.concatenate([X[c].to_numpy(dtype=np.float64) for c in cols]) if n_cov else np.zeros(0)
    else:                                     # long: (variable, value) pairs
        variables = list(pd.unique(X[variable_column_name]))
        vcode = {v: i for i, v in enumerate(variables)}
        n_cov = len(variables)
        s_long, t_long = s_code, t
        v_long = X[variable_column_name].map(vcode).to_numpy(dtype=np.int64)
        x_long = X[value_column_name].to_numpy(dtype=np.float64)
    keep = ~np.isnan(s_long) if s_long.dtype.kind == "f" else np.ones(s_long.shape[0], dtype=bool)    # rows of unknown series
    s_long, v_long, t_long, x_long = s_long[keep].astype(np.int64), v_long[keep], t_long[keep], x_long[keep]
    order = np.lexsort((t_long, v_long, s_long))              # by series, then variable, then time (stable)
    cnt = np.bincount(s_long * n_cov + v_long, minlength=n_series * n_cov) if n_cov else np.zeros(0, dtype=np.int64)
    per_series = cnt.reshape(n_series, n_cov) if n_cov else cnt
    if n_cov and np.any((per_series == 0).any(axis=1) & (per_series > 0).any(axis=1)):
        raise ValueError("every series needs the same covariates")
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return n_cov, off, t_long[order], x_long[order]


def one_hot_series_covariates(data: pd.DataFrame, time_column_name="time", series_column_name="series", distance=1.0) -> pd.DataFrame:
    """Covariate frame that hands the network a one-hot encoding of the series (scaled by `distance`): the reference's
    `MultiTimeSeriesNetwork` (src/NeuralNetworkConstructors.jl:60-82) and the `vcat(u, one_hot)` model of
    docs/src/MultipleTimeSeries.md:72-92.  One row per series: a single knot is a constant covariate (flat extrapolation,
    src/covariates.jl:45-48), so `NN([u; X(t, i)])` IS `NN([u; distance * onehot(i)])`.  Use with

        X = one_hot_series_covariates(data)
        rhs = expression_rhs(d=2, n_cov=X.shape[1] - 2, nn_out=2, du=["y1", "y2"])        # any m: kernels are generated on demand
        model = MultiCustomDerivatives(data, rhs, {"NN": SimpleNeuralNetwork(2 + m, 2)[1]}, X=X)

    (for m <= 2 series the catalog's `MultiNODE(data, X=X)` has pre-compiled kernels)."""
    labels = sorted(pd.unique(data[series_column_name]))
    t0 = data.groupby(series_column_name)[time_column_name].min()
    rows = []
    for i, lab in enumerate(labels):
        row = {time_column_name: float(t0[lab]), series_column_name: lab}
        row.update({f"series_is_{l2}": (float(distance) if j == i else 0.0) for j, l2 in enumerate(labels)})
        rows.append(row)
    return pd.DataFrame(rows)


# ---- model objects ---------------------------------------------------------------------------------------------------
@dataclass
class UDE:
    """Mirror of the reference's UDE / MultiUDE structs (src/Models.jl:26-46, src/MultipleTimeSeries.jl:27-49)."""
    times: np.ndarray
    data: np.ndarray                      # d x N
    data_frame: pd.DataFrame
    X_data_frame: Optional[pd.DataFrame]
    rhs: Rhs
    hidden: int
    parameters: np.ndarray                # flat theta = [uhat(:) | process_model]
    pm_names: List[str]
    weights: Dict[str, float]             # regularization, process, observation
    constructor: Callable
    time_column_name: str
    solvers: object
    varnames: List[str]
    N: int = 0
    T: float = 0.0
    multi: bool = False
    series_column_name: Optional[str] = None
    starts: Optional[np.ndarray] = None
    lengths: Optional[np.ndarray] = None
    labels: Optional[list] = None
    cov: Optional[tuple] = None           # (n_cov, knot_off, knot_t, knot_x)
    obs_divisor: float = 1.0              # ObservationMSE(N | T, w): MultiNODE divides by T (App. A.3)
    uhat_init: float = 1e-3
    extra: dict = field(default_factory=dict)

    # -- layout -------------------------------------------------------------------------------------------------------
    @property
    def d(self):
        return self.data.shape[0]

    @property
    def n_cov(self):
        return self.cov[0] if self.cov else 0

    def base_problem(self) -> P.Problem:
        """Model part of the Problem (no loss yet), with the reference's parameter order: NN leaves, then scalars."""
        n_series = 1 if not self.multi else len(self.starts)
        p = P.Problem(d=self.d, nn_in=self.rhs.nn_in, nn_hidden=self.hidden, nn_out=self.rhs.nn_out, rhs_kind=self.rhs.kind,
                      n_cov=self.n_cov, n_scalars=len(self.rhs.scalar_names), has_series_param=self.rhs.series_param is not None,
                      solver=self.solvers.kind, substeps=self.solvers.substeps, time_scale=self.rhs.time_scale,
                      time_shift=self.rhs.time_shift)
        if self.rhs.kind >= 1000:           # expression right-hand side (user_rhs.expression_rhs)
            p.user_rhs = self.rhs
        p.default_layout(n_series_param=n_series)
        p.data, p.times = self.data, self.times
        if self.multi:
            p.starts, p.lengths = self.starts, self.lengths
        else:
            p.starts, p.lengths = np.array([0]), np.array([self.data.shape[1]])
        if self.cov:
            _, p.knot_off, p.knot_t, p.knot_x = self.cov
        return p

    @property
    def uhat(self):
        return self.parameters[: self.d * self.data.shape[1]].reshape(self.d, -1, order="F")

    @property
    def process_model(self):
        return self.parameters[self.d * self.data.shape[1]:]


def _pack_pm(prob: P.Problem, rhs: Rhs, initial_parameters: dict, n_series: int) -> np.ndarray:
    nn = initial_parameters.get("NN") or initial_parameters.get("NNparams") or initial_parameters.get("NNparameters")
    if nn is None:
        raise ValueError("initial_parameters needs an `NN` entry (src/helpers.jl:506-513 aliases)")
    pm = np.zeros(prob.n_pm)
    H, din, dout = prob.nn_hidden, prob.nn_in, prob.nn_out
    W1, W2 = np.asarray(nn["layer_1.weight"], dtype=np.float64), np.asarray(nn["layer_2.weight"], dtype=np.float64)
    if W1.shape != (H, din) or W2.shape != (dout, H):
        raise ValueError(f"NN shape {W1.shape}->{W2.shape} does not match the right-hand side ({din}->{H}->{dout})")
    pm[prob.off_w1:prob.off_w1 + H * din] = W1.reshape(-1, order="F")
    pm[prob.off_b1:prob.off_b1 + H] = nn["layer_1.bias"]
    pm[prob.off_w2:prob.off_w2 + dout * H] = W2.reshape(-1, order="F")
    pm[prob.off_b2:prob.off_b2 + dout] = nn["layer_2.bias"]
    for o, nm in zip(prob.off_scalar, rhs.scalar_names):
        if nm.startswith("beta") and "beta" in initial_parameters:
            pm[o] = float(np.asarray(initial_parameters["beta"])[int(nm[4:]) - 1])
        else:
            pm[o] = float(initial_parameters[nm])
    if rhs.series_param is not None:
        v = np.asarray(initial_parameters[rhs.series_param], dtype=np.float64)
        if v.shape[0] < n_series:
            raise ValueError("per-series parameter vector must have one entry per series")
        # a longer vector is indexed by the position of the series among those present, as `p.r[round(Int, i)]` is in the
        # reference when a CV driver rebuilds the model on fewer series (src/CrossValidation.jl:463-474)
        pm[prob.off_series_param:prob.off_series_param + n_series] = v[:n_series]
    return pm


def _finish(model: UDE, initial_parameters: dict) -> UDE:
    prob = model.base_problem()
    n_series = 1 if not model.multi else len(model.starts)
    pm = _pack_pm(prob, model.rhs, initial_parameters, n_series)
    uhat = np.full(model.d * model.data.shape[1], model.uhat_init)
    model.parameters = np.concatenate([uhat, pm])
    return model


def CustomDerivatives(data, derivs: Rhs, initial_parameters: dict, X=None, time_column_name="time", variable_column_name=None,
                      value_column_name=None, proc_weight=1.0, obs_weight=1.0, reg_weight=1e-6, extrap_rho=0.1, l=0.25,
                      reg_type="L2", ode_solver=CudaRK4(), ad_method="discrete adjoint"):
    """src/Models.jl:115-155 (and :248-288 with covariates).  `derivs` is an RHS-catalog entry."""
    if reg_type != "L2":
        raise ValueError("L1 regularisation throws in the reference (abs on an array, src/Regularization.jl:50); use L2")
    N, dims, T, times, mat, df, varnames = process_data(data, time_column_name)
    cov = interpolate_covariates(X, time_column_name, None, None, variable_column_name, value_column_name) if X is not None else None
    hidden = np.asarray(initial_parameters["NN"]["layer_1.bias"]).shape[0]
    ctor = lambda d2: CustomDerivatives(d2, derivs, initial_parameters, X, time_column_name, variable_column_name, value_column_name,
                                        proc_weight, obs_weight, reg_weight, extrap_rho, l, reg_type, ode_solver, ad_method)
    m = UDE(times, mat, df, X, derivs, hidden, np.zeros(0), [], dict(regularization=reg_weight, process=proc_weight, observation=obs_weight),
            ctor, time_column_name, ode_solver, varnames, N=N, T=T, cov=cov, obs_divisor=float(N))

    def reseed(d2, sd):
        p2 = dict(initial_parameters)
        p2["NN"] = SimpleNeuralNetwork(derivs.nn_in, derivs.nn_out, hidden=hidden, seed=sd)[1]
        return CustomDerivatives(d2, derivs, p2, X, time_column_name, variable_column_name, value_column_name, proc_weight, obs_weight,
                                 reg_weight, extrap_rho, l, reg_type, ode_solver, ad_method)
    m.extra.update(l=l, extrap_rho=extrap_rho, reseed=reseed)
    return _finish(m, initial_parameters)


def NODE(data, X=None, time_column_name="time", variable_column_name=None, value_column_name=None, hidden_units=10, seed=1,
         proc_weight=1.0, obs_weight=1.0, reg_weight=1e-6, reg_type="L2", l=0.25, extrap_rho=0.1, ode_solver=CudaRK4(),
         ad_method="discrete adjoint"):
    """src/Models.jl:746-840: NN(vcat(u, covariates(t))) with hidden_units tanh units."""
    N, dims, T, times, mat, df, varnames = process_data(data, time_column_name)
    cov = interpolate_covariates(X, time_column_name, None, None, variable_column_name, value_column_name) if X is not None else None
    n_cov = cov[0] if cov else 0
    _, nn = SimpleNeuralNetwork(dims + n_cov, dims, hidden=hidden_units, seed=seed)
    ctor = lambda d2: NODE(d2, X, time_column_name, variable_column_name, value_column_name, hidden_units, seed, proc_weight,
                           obs_weight, reg_weight, reg_type, l, extrap_rho, ode_solver, ad_method)
    m = UDE(times, mat, df, X, node_rhs(dims, n_cov), hidden_units, np.zeros(0), [],
            dict(regularization=reg_weight, process=proc_weight, observation=obs_weight), ctor, time_column_name, ode_solver,
            varnames, N=N, T=T, cov=cov, obs_divisor=float(N))
    m.extra.update(l=l, extrap_rho=extrap_rho,
                   reseed=lambda d2, sd: NODE(d2, X, time_column_name, variable_column_name, value_column_name, hidden_units, sd,
                                              proc_weight, obs_weight, reg_weight, reg_type, l, extrap_rho, ode_solver, ad_method))
    return _finish(m, {"NN": nn})


def _multi(data, X, derivs, initial_parameters, time_column_name, series_column_name, variable_column_name, value_column_name,
           proc_weight, obs_weight, reg_weight, ode_solver, hidden, ctor, obs_div_T, uhat_init):
    N, T, dims, mat, times, df, starts, lengths, varnames, labels = process_multi_data(data, time_column_name, series_column_name)
    cov = interpolate_covariates(X, time_column_name, series_column_name, labels, variable_column_name, value_column_name) if X is not None else None
    rhs = derivs(dims, cov[0] if cov else 0) if callable(derivs) else derivs
    m = UDE(times, mat, df, X, rhs, hidden, np.zeros(0), [], dict(regularization=reg_weight, process=proc_weight, observation=obs_weight),
            ctor, time_column_name, ode_solver, varnames, N=N, T=T, multi=True, series_column_name=series_column_name,
            starts=starts, lengths=lengths, labels=labels, cov=cov, obs_divisor=(T if obs_div_T else float(N)), uhat_init=uhat_init)
    return m


def MultiNODE(data, X=None, time_column_name="time", series_column_name="series", variable_column_name=None, value_column_name=None,
              hidden_units=10, seed=1, proc_weight=1.0, obs_weight=1.0, reg_weight=1e-6, reg_type="L2", l=0.5, extrap_rho=0.0,
              ode_solver=CudaRK4(), ad_method="discrete adjoint"):
    """src/MultipleTimeSeries.jl:181-279.  Quirks kept: uhat starts at 0 (:200) and ObservationMSE divides by T (:190)."""
    ctor = lambda d2: MultiNODE(d2, X, time_column_name, series_column_name, variable_column_name, value_column_name, hidden_units,
                                seed, proc_weight, obs_weight, reg_weight, reg_type, l, extrap_rho, ode_solver, ad_method)
    m = _multi(data, X, node_rhs, None, time_column_name, series_column_name, variable_column_name, value_column_name, proc_weight,
               obs_weight, reg_weight, ode_solver, hidden_units, ctor, True, 0.0)
    _, nn = SimpleNeuralNetwork(m.rhs.nn_in, m.rhs.nn_out, hidden=hidden_units, seed=seed)
    m.extra.update(l=l, extrap_rho=extrap_rho,
                   reseed=lambda d2, sd: MultiNODE(d2, X, time_column_name, series_column_name, variable_column_name, value_column_name,
                                                   hidden_units, sd, proc_weight, obs_weight, reg_weight, reg_type, l, extrap_rho,
                                                   ode_solver, ad_method))
    return _finish(m, {"NN": nn})


def MultiCustomDerivatives(data, derivs: Rhs, initial_parameters: dict, X=None, time_column_name="time", series_column_name="series",
                           variable_column_name=None, value_column_name=None, proc_weight=1.0, obs_weight=1.0, reg_weight=1e-6,
                           reg_type="L2", extrap_rho=0.1, l=0.25, ode_solver=CudaRK4(), ad_method="discrete adjoint"):
    """src/MultipleTimeSeries.jl:301-362."""
    hidden = np.asarray(initial_parameters["NN"]["layer_1.bias"]).shape[0]
    ctor = lambda d2: MultiCustomDerivatives(d2, derivs, initial_parameters, X, time_column_name, series_column_name, variable_column_name,
                                             value_column_name, proc_weight, obs_weight, reg_weight, reg_type, extrap_rho, l, ode_solver, ad_method)
    m = _multi(data, X, derivs, initial_parameters, time_column_name, series_column_name, variable_column_name, value_column_name,
               proc_weight, obs_weight, reg_weight, ode_solver, hidden, ctor, False, 1e-3)
    return _finish(m, initial_parameters)
