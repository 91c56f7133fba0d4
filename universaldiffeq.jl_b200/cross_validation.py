"""Cross-validation drivers of the reference, on the CUDA hot path.

  leave_future_out / leave_future_out_cv   /root/reference/src/CrossValidation.jl:17-45, :128-141 (UDE), :234-278, :354-366 (MultiUDE)
  summarize_leave_future_out               :91-110
  interval_holdout_cv                      single-interval hold-outs via t_skip (the reference's internal kfold_cv, :167-195, is out of scope)
The reference trains the k folds under Threads.@threads, one optimiser per fold.  Here the folds are the *model* batch
axis of one handle (ude_set_data's model_of_series): `leave_future_out_batched` trains all folds x ensemble members in
one device-resident Adam run -- the folds are independent theta vectors, Adam is element-wise, so this is exactly k
separate train! calls.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np
import pandas as pd

from . import problem as P
from ._capi import Engine
from .models import UDE
from .training import build_problem, default_options, forecast


def _split(model: UDE, k: int, skip: int):
    data = model.data_frame
    tcol = model.time_column_name
    train, test = [], []
    for i in range(skip, skip * k + 1, skip):
        if not model.multi:
            train.append(data.iloc[:-i]); test.append(data.iloc[-i:])
        else:   # src/CrossValidation.jl:242-255: drop the last i points of every series (one sort, not one filter per series)
            srt = data.sort_values([model.series_column_name, tcol], kind="stable")
            from_end = srt.groupby(model.series_column_name, sort=False).cumcount(ascending=False).to_numpy()
            train.append(srt[from_end >= i].reset_index(drop=True))
            test.append(srt[from_end < i].reset_index(drop=True))
    return train, test


def _forecast_fold(model_i: UDE, test_i: pd.DataFrame, device=0) -> pd.DataFrame:
    tcol = model_i.time_column_name
    if not model_i.multi:
        tt = np.sort(test_i[tcol].to_numpy(dtype=np.float64))
        out = forecast(model_i, model_i.uhat[:, -1], model_i.times[-1], tt, device=device)
        df = pd.DataFrame(out, columns=[tcol] + model_i.varnames)
        df["horizon"] = df[tcol] - model_i.times.max()
        return df
    frames = []
    for s, lab in enumerate(model_i.labels):
        tt = np.sort(test_i[test_i[model_i.series_column_name] == lab][tcol].to_numpy(dtype=np.float64))
        if tt.size == 0:
            continue
        c_last = int(model_i.starts[s] + model_i.lengths[s] - 1)
        out = forecast(model_i, model_i.uhat[:, c_last], model_i.times[c_last], tt, series=s, device=device)
        df = pd.DataFrame(out, columns=[tcol] + model_i.varnames)
        df[model_i.series_column_name] = lab
        df["horizon"] = df[tcol] - model_i.times[c_last]
        frames.append(df)
    return pd.concat(frames, ignore_index=True)


def leave_future_out(model: UDE, training: Callable[[UDE], None], k: int, skip: int = 1, device=0):
    """leave_future_out(model, training!, k; skip): k refits on data[1:end-i], forecasts of the held-out tail.
    Returns (training_data, testing_data, forecasts) like leave_future_out_cv (src/CrossValidation.jl:17-45)."""
    train, test = _split(model, k, int(round(skip)))
    forecasts = []
    for tr, te in zip(train, test):
        m = model.constructor(tr)           # same seed => same NN init in every fold
        training(m)
        forecasts.append(_forecast_fold(m, te, device))
    return train, test, forecasts


def summarize_leave_future_out(model: UDE, testing: List[pd.DataFrame], forecasts: List[pd.DataFrame]):
    """Mean absolute error by forecast horizon (src/CrossValidation.jl:91-110)."""
    rows = []
    keys = [model.time_column_name] + ([model.series_column_name] if model.multi else [])
    for te, fc in zip(testing, forecasts):
        mrg = fc.merge(te, on=keys, suffixes=("_forecast", "_testing"))
        for v in model.varnames:
            for h, a, b in zip(mrg["horizon"], mrg[v + "_forecast"], mrg[v + "_testing"]):
                rows.append((h, v, abs(a - b)))
    df = pd.DataFrame(rows, columns=["horizon", "variable", "abs_error"])
    by_h = df.groupby("horizon")["abs_error"].agg(mean_absolute_error="mean", standard_error_of_MAE=lambda x: x.std(ddof=1) / np.sqrt(len(x)) if len(x) > 1 else 0.0)
    return by_h.reset_index(), df


def stack_models(models: List[UDE], loss_function, regularization_weight=0.0, loss_options=None) -> tuple:
    """One Problem whose model axis holds `models` (folds and/or ensemble members of the SAME architecture):
    returns (problem, theta_all).  Per-model loss scales keep each fold's own N and T (src/LossFunctions.jl:17-28)."""
    probs = [build_problem(m, loss_function, regularization_weight, loss_options) for m in models]
    p0 = probs[0]
    big = P.Problem(d=p0.d, nn_in=p0.nn_in, nn_hidden=p0.nn_hidden, nn_out=p0.nn_out, rhs_kind=p0.rhs_kind, n_cov=p0.n_cov,
                    n_scalars=p0.n_scalars, has_series_param=p0.has_series_param, solver=p0.solver, substeps=p0.substeps,
                    time_scale=p0.time_scale, time_shift=p0.time_shift, loss_kind=p0.loss_kind, pred_length=p0.pred_length,
                    proc_inv_dt=p0.proc_inv_dt, shoot_from_theta=p0.shoot_from_theta, shoot_first_scored=p0.shoot_first_scored,
                    remove_ends=p0.remove_ends, preroll_eps=p0.preroll_eps, A=p0.A, B=p0.B)
    big.user_rhs = p0.user_rhs
    big.off_w1, big.off_b1, big.off_w2, big.off_b2 = p0.off_w1, p0.off_b1, p0.off_w2, p0.off_b2
    big.off_scalar, big.off_series_param, big.n_pm = list(p0.off_scalar), p0.off_series_param, p0.n_pm
    col, ser = 0, 0
    starts, lengths, mos, knot_off, kt, kx = [], [], [], [0], [], []
    for mi, p in enumerate(probs):
        if (p.n_pm, p.d, p.rhs_kind) != (p0.n_pm, p0.d, p0.rhs_kind):
            raise ValueError("stacked models must share the architecture")
        starts.append(p.starts + col); lengths.append(p.lengths); mos.append(np.full(p.n_series, mi))
        if p.n_cov:
            knot_off.extend((p.knot_off[1:] + knot_off[-1]).tolist()); kt.append(p.knot_t); kx.append(p.knot_x)
        col += p.n_cols; ser += p.n_series
    big.starts, big.lengths, big.model_of_series = np.concatenate(starts), np.concatenate(lengths), np.concatenate(mos)
    big.times = np.concatenate([p.times for p in probs])
    big.data = np.concatenate([np.asarray(p.data) for p in probs], axis=1)
    for name in ("obs_scale", "proc_scale", "shoot_weight", "reg_weight"):
        setattr(big, name, np.array([getattr(p, name)[0] for p in probs]))
    if p0.n_cov:
        big.knot_off, big.knot_t, big.knot_x = np.array(knot_off), np.concatenate(kt), np.concatenate(kx)
    if p0.loss_kind == P.LOSS_DERIV_MATCH:
        big.dm_u = np.concatenate([np.asarray(p.dm_u) for p in probs], axis=1)
        big.dm_dudt = np.concatenate([np.asarray(p.dm_dudt) for p in probs], axis=1)
    if any(p.skip_mask is not None for p in probs):
        big.skip_mask = np.concatenate([p.skip_mask if p.skip_mask is not None else np.zeros(p.n_cols, np.uint8) for p in probs])
    big.finalize()
    theta = np.concatenate([m.parameters for m in models])
    assert theta.shape[0] == big.n_theta
    return big, theta


def leave_future_out_batched(model: UDE, k: int, skip: int = 1, loss_function="conditional likelihood", regularization_weight=0.0,
                             loss_options=None, optim_options=None, n_members: int = 1, member_seeds: Optional[List[int]] = None,
                             device=0, do_forecast=True, devices=None):
    """All k folds x n_members ensemble members (same model re-initialised with different NN seeds through
    `model.extra['reseed'](data, seed)`; BASELINE config C4) trained in ONE device-resident Adam run: member-major,
    fold-minor.  devices=(0, 1, ...) spreads the independent models over several GPUs (whole models per GPU, no exchange).
    Returns (models, training_data, testing_data, forecasts, loss_trace)."""
    train, test = _split(model, k, int(round(skip)))
    if n_members > 1 or member_seeds is not None:
        seeds = list(member_seeds) if member_seeds is not None else list(range(1, n_members + 1))
        if "reseed" not in model.extra:
            raise ValueError("this model cannot be re-seeded")
        models = [model.extra["reseed"](tr, sd) for sd in seeds for tr in train]
        test = [te for _ in seeds for te in test]
        train = [tr for _ in seeds for tr in train]
    else:
        models = [model.constructor(tr) for tr in train]
    if loss_function == "marginal likelihood":       # UKF loss: all folds x members filtered in one many-model handle
        from .training import marginal_likelihood_many_
        marginal_likelihood_many_(models, regularization_weight, loss_options, optim_options, device=device)
        trace = np.stack([m.extra["loss_trace"] for m in models], axis=1)
        forecasts = [_forecast_fold(m, te, device) for m, te in zip(models, test)] if do_forecast else []
        return models, train, test, forecasts, trace
    opts = dict(default_options(loss_function)); opts.update(optim_options or {})
    big, theta = stack_models(models, loss_function, regularization_weight, loss_options)
    if devices is not None and len(devices) > 1 and big.n_models >= len(devices):
        from .sharding import shard_problem
        from .training import _run_on_ranks
        world = len(devices)
        shards = [shard_problem(big, theta, r, world) for r in range(world)]

        def fit(r):
            with Engine(shards[r].problem, device=devices[r]) as eng:
                return eng.adam(shards[r].theta, opts["step_size"], opts["maxiter"])

        res = _run_on_ranks([lambda r=r: fit(r) for r in range(world)])
        theta = theta.copy()
        for sh, (th_r, _) in zip(shards, res):
            lo, hi, _ = sh.uhat_slices[0]
            theta[lo:hi] = th_r
        trace = np.concatenate([tr for _, tr in res], axis=1)
    else:
        with Engine(big, device=device) as eng:
            theta, trace = eng.adam(theta, opts["step_size"], opts["maxiter"])
    for mi, m in enumerate(models):
        lo = int(big.theta_base[mi]); hi = int(big.theta_base[mi + 1]) if mi + 1 < big.n_models else big.n_theta
        m.parameters = theta[lo:hi].copy()
    forecasts = [_forecast_fold(m, te, device) for m, te in zip(models, test)] if do_forecast else []
    return models, train, test, forecasts, trace


def leave_future_out_predict(model: UDE, training: Callable[[UDE], None], n_per_fold: int, k_folds: int, device=0):
    """leave_future_out_predict(model, training!, n_per_fold, k_folds) (src/CrossValidation.jl:608-640): fold i trains on
    data[1:end-i*n_per_fold] and is scored on one-step-ahead CHANGES over the following n_per_fold intervals (the last
    training-side point is included as the first start).  Returns (summary, raw) like summarize_cv_results (:525-564):
    per variable MSE, MSE_SE, MAE, MAE_SE, correlation, count; raw has time, variable, observed, predicted, AE, SE."""
    from .training import predictions
    if model.multi:
        raise NotImplementedError("leave_future_out_predict is defined for single-series models in the reference")
    data = model.data_frame.sort_values(model.time_column_name).reset_index(drop=True)
    n = len(data)
    rows = []
    for i in range(n_per_fold, n_per_fold * k_folds + 1, n_per_fold):
        train_i, test_i = data.iloc[: n - i], data.iloc[n - i - 1: n - i + n_per_fold]
        m = model.constructor(train_i)
        training(m)
        inits, obs, preds = predictions(m, test_i, device=device)
        tt = np.sort(test_i[model.time_column_name].to_numpy(dtype=np.float64))[1:]
        for v, name in enumerate(model.varnames):
            for j in range(tt.shape[0]):
                rows.append((tt[j], name, obs[v, j] - inits[v, j], preds[v, j] - inits[v, j]))
    raw = pd.DataFrame(rows, columns=[model.time_column_name, "variable", "observed", "predicted"])
    raw["AE"] = (raw.observed - raw.predicted).abs()
    raw["SE"] = (raw.observed - raw.predicted) ** 2
    se = lambda x: x.std(ddof=1) / np.sqrt(len(x)) if len(x) > 1 else 0.0          # noqa: E731
    out = []
    for name, g in raw.groupby("variable", sort=False):
        cor = np.corrcoef(g.observed, g.predicted)[0, 1] if len(g) > 1 else np.nan
        out.append((name, g.SE.mean(), se(g.SE), g.AE.mean(), se(g.AE), cor, len(g)))
    summary = pd.DataFrame(out, columns=["variable", "MSE", "MSE_SE", "MAE", "MAE_SE", "correlation", "count"])
    return summary, raw


def predict(model: UDE, test_data: pd.DataFrame, device=0) -> pd.DataFrame:
    """predict(UDE, test_data) (src/MultiModelTests.jl:124-150; single series: src/ModelTesting.jl:245-260): one-step-ahead
    predictions between consecutive rows of `test_data`, started from the observed values; for multi-series models each
    series of the test frame is predicted with the covariates of the TRAINED series that carries the same label.  All
    intervals of a series are one batch of segments on the GPU.  Returns [series,] time, variables at the interval ends."""
    from .training import _interval_problem, _interval_pm
    tcol, scol = model.time_column_name, model.series_column_name
    frames = []
    groups = [(None, test_data)] if not model.multi else [(lab, test_data[test_data[scol] == lab]) for lab in sorted(pd.unique(test_data[scol]))]
    for lab, df in groups:
        df = df.sort_values(tcol)
        j = 0 if lab is None else list(model.labels).index(lab)
        tt = df[tcol].to_numpy(dtype=np.float64)
        Y = df[list(model.varnames)].to_numpy(dtype=np.float64).T
        if tt.shape[0] < 2:
            continue
        pr = _interval_problem(model, tt[:-1], tt[1:], j).finalize()
        cols = np.zeros((model.d, 2 * (tt.shape[0] - 1)), order="F")
        cols[:, 0::2] = Y[:, :-1]
        with Engine(pr, device=device) as eng:
            pred = eng.forward(np.concatenate([cols.reshape(-1, order="F"), _interval_pm(model, pr, j)]))[:, 1::2]
        out = pd.DataFrame(pred.T, columns=list(model.varnames))
        out.insert(0, tcol, tt[1:])
        if lab is not None:
            out.insert(0, scol, lab)
        frames.append(out)
    return pd.concat(frames, ignore_index=True)


def energy_score(testing: pd.DataFrame, preds: List[pd.DataFrame], model: UDE) -> float:
    """src/CrossValidation.jl:422-447: mean over test rows of sum_i [ |pred_i - y|_1 / M - sum_j |pred_i - pred_j|_1 / (2 M^2) ]."""
    Y = testing[list(model.varnames)].to_numpy(dtype=np.float64)
    Pm = np.stack([p[list(model.varnames)].to_numpy(dtype=np.float64) for p in preds])          # M x T x d
    M = Pm.shape[0]
    a = np.abs(Pm - Y[None]).sum(axis=2).sum(axis=0) / M
    b = np.abs(Pm[:, None] - Pm[None, :]).sum(axis=3).sum(axis=(0, 1)) / (2.0 * M * M)
    return float(np.mean(a - b))


def leave_site_out(model: UDE, training: Callable[[UDE], None], sites: Optional[int] = None, seed: int = 0, device=0) -> float:
    """leave_site_out(model, training!; sites) (src/CrossValidation.jl:463-474): for each left-out site, refit on the others and
    score the one-step-ahead predictions of the left-out series made with every remaining site's covariates (energy score)."""
    if not model.multi:
        raise ValueError("leave_site_out needs a multi-series model")
    scol, tcol = model.series_column_name, model.time_column_name
    labels = list(pd.unique(model.data_frame[scol]))
    n = len(labels) if sites is None else int(sites)
    chosen = list(np.random.default_rng(seed).choice(np.array(labels, dtype=object), size=n, replace=False))
    scores = []
    for lab in chosen:
        train_i = model.data_frame[model.data_frame[scol] != lab]
        test_i = model.data_frame[model.data_frame[scol] == lab].copy()
        m = model.constructor(train_i)
        training(m)
        preds = []
        for j in m.labels:
            test_i[scol] = j
            preds.append(predict(m, test_i, device=device))
        tmin = test_i[tcol].min()
        scores.append(energy_score(test_i[test_i[tcol] > tmin].sort_values(tcol), preds, m))
    return float(np.mean(scores))


def interval_holdout_cv(model: UDE, training: Callable[[UDE, float], None], k: int = 10):
    """k single-interval hold-outs through the t_skip variant of the conditional likelihood (LFC.jl:68-95): `training(model,
    t_skip)` must fit with that interval left out; returns the one-step prediction errors at the skipped intervals.
    NOT the reference's internal `kfold_cv` (src/CrossValidation.jl:167-195: contiguous block folds, refit on the concatenated
    rest, scored by the UKF marginal likelihood of the held-out block) -- that driver is outside the hot-path scope
    ("CV summaries"); its two ingredients are here: the t_skip loss and `Engine.ukf_loss_grad`."""
    N = model.data.shape[1]
    idx = np.unique(np.linspace(2, N - 1, k).round().astype(int))      # 1-based END index of the interval left out
    rows = []
    from .training import predictions
    for t in idx:
        m = model.constructor(model.data_frame)
        training(m, float(t))
        pred = predictions(m)[:, t - 1]
        rows.append((t, float(np.mean(np.abs(pred - m.data[:, t - 1])))))
    return pd.DataFrame(rows, columns=["t_skip", "mean_absolute_error"])
