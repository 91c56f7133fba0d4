"""Data-parallel sharding of one Problem across the GPUs of a box.

Single-model problems (MultiNODE / MultiUDE fits) are split into contiguous blocks of *series*: a series' columns,
uhat block, covariate knots and all of its segments stay on one rank; the process-model gradient and the loss are the
only quantities exchanged (one sum-allreduce of n_shared + 1 doubles per evaluation, done by the library over NCCL).
Many-model problems (CV folds x ensemble members) are split by *model*: no exchange at all.
Pure NumPy; tests/test_sharding_gloo.py exercises it with world_size 2 on CPU.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass

import numpy as np

from . import problem as P


@dataclass
class Shard:
    problem: P.Problem
    theta: np.ndarray
    uhat_slices: list        # [(global_lo, global_hi, local_lo)] ranges of theta_global copied into theta_local (uhat blocks)
    pm_shared: int           # leading entries of process_model that are shared across ranks (allreduced)
    series: tuple            # (first, last+1) global series owned
    models: tuple            # (first, last+1) global models owned


def _bounds(n, world, rank):
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_problem(prob: P.Problem, theta: np.ndarray, rank: int, world: int) -> Shard:
    if world == 1:
        return Shard(prob, theta, [(0, prob.n_theta, 0)], prob.n_pm if not prob.has_series_param else prob.off_series_param,
                     (0, prob.n_series), (0, prob.n_models))
    d = prob.d
    by_model = prob.n_models > 1
    if by_model:
        m0, m1 = _bounds(prob.n_models, world, rank)
        s0, s1 = int(prob.model_series0[m0]), int(prob.model_series0[m1])
    else:
        m0, m1 = 0, 1
        s0, s1 = _bounds(prob.n_series, world, rank)
    if s1 <= s0:
        raise ValueError("more ranks than series/models")
    c0, c1 = int(prob.starts[s0]), int(prob.starts[s1 - 1] + prob.lengths[s1 - 1])
    q = copy.copy(prob)
    q.starts = prob.starts[s0:s1] - c0
    q.lengths = prob.lengths[s0:s1].copy()
    q.times = prob.times[c0:c1].copy()
    q.data = np.asfortranarray(prob.data[:, c0:c1])
    q.model_of_series = (prob.model_of_series[s0:s1] - m0) if by_model else None
    for name in ("obs_scale", "proc_scale", "shoot_weight", "reg_weight"):
        setattr(q, name, getattr(prob, name)[m0:m1].copy())
    if not by_model and rank != 0:
        q.reg_weight = np.zeros(1)             # the regulariser depends on shared parameters only: rank 0 carries it
    if prob.n_cov:
        k0, k1 = int(prob.knot_off[s0 * prob.n_cov]), int(prob.knot_off[s1 * prob.n_cov])
        q.knot_off = prob.knot_off[s0 * prob.n_cov: s1 * prob.n_cov + 1] - k0
        q.knot_t, q.knot_x = prob.knot_t[k0:k1].copy(), prob.knot_x[k0:k1].copy()
    if prob.loss_kind == P.LOSS_DERIV_MATCH:
        q.dm_u, q.dm_dudt = np.asfortranarray(prob.dm_u[:, c0:c1]), np.asfortranarray(prob.dm_dudt[:, c0:c1])
    if prob.skip_mask is not None:
        q.skip_mask = prob.skip_mask[c0:c1].copy()
    pm_shared = prob.n_pm
    if prob.has_series_param:
        if by_model:
            pass                                   # whole models move: their series vectors move with them
        else:
            if prob.off_series_param + prob.n_series != prob.n_pm:
                raise ValueError("the per-series vector must be the last entry of process_model to be sharded")
            pm_shared = prob.off_series_param
            q.n_pm = prob.off_series_param + (s1 - s0)
    q.finalize()
    th = np.zeros(q.n_theta)
    slices = []
    if by_model:
        lo = int(prob.theta_base[m0]); hi = int(prob.theta_base[m1]) if m1 < prob.n_models else prob.n_theta
        th[:] = theta[lo:hi]
        slices.append((lo, hi, 0))
    else:
        th[: d * (c1 - c0)] = theta[d * c0: d * c1]
        slices.append((d * c0, d * c1, 0))
        pmg = prob.pm_base(0)
        th[d * (c1 - c0): d * (c1 - c0) + pm_shared] = theta[pmg: pmg + pm_shared]
        if prob.has_series_param:
            th[d * (c1 - c0) + pm_shared:] = theta[pmg + pm_shared + s0: pmg + pm_shared + s1]
    return Shard(q, th, slices, pm_shared, (s0, s1), (m0, m1))


def gather_gradient(prob: P.Problem, shards, local_grads, reduced_pm_grad=None) -> np.ndarray:
    """Reassemble the global gradient from per-rank local gradients (host-side check / single-process multi-GPU).
    `reduced_pm_grad`: the already allreduced shared process-model gradient (else it is summed here)."""
    g = np.zeros(prob.n_theta)
    d = prob.d
    if prob.n_models > 1:
        for sh, gl in zip(shards, local_grads):
            lo, hi, _ = sh.uhat_slices[0]
            g[lo:hi] = gl
        return g
    pmg = prob.pm_base(0)
    acc = np.zeros(shards[0].pm_shared)
    for sh, gl in zip(shards, local_grads):
        lo, hi, _ = sh.uhat_slices[0]
        n_u = hi - lo
        g[lo:hi] = gl[:n_u]
        acc += gl[n_u:n_u + sh.pm_shared]
        if prob.has_series_param:
            s0, s1 = sh.series
            g[pmg + sh.pm_shared + s0: pmg + sh.pm_shared + s1] = gl[n_u + sh.pm_shared:]
    g[pmg:pmg + shards[0].pm_shared] = acc if reduced_pm_grad is None else reduced_pm_grad
    return g
