"""Third, independent gradient of the fixed-step losses: torch.float64 reverse-mode autograd.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (no Julia in this image; see oracle/ude_oracle.h).  The reference differentiates its losses with Zygote
through the ODE solve (src/Optimizers.jl:476, AutoZygote); with a FIXED step sequence that is the derivative of the discrete
map, which is what this file asks torch.autograd for.  It shares no code with ude_oracle.c (hand-derived adjoint), with
ude_oracle_np.py (complex step) or with the CUDA kernels: the losses are re-assembled here from the reference's formulas
  conditional likelihood   LFC.jl:17-53, :68-95 (t_skip), multi :347-404
  default MSE loss         src/helpers.jl:18-44, src/LossFunctions.jl:17-28
  multiple shooting        LFC.jl:261-305 (single, 1e-6 pre-roll), :639-708 (multi)
  shooting                 LFC.jl:196-232, :545-601
  derivative matching      LFC.jl:175-192, :500-540
  L2 on the weight leaves  src/Regularization.jl:19-36, :97-105
and the right-hand-side catalog of SURVEY.md App. G.  Whole series are advanced as batched tensors where the loss allows
it (every interval of a state-space loss at once), so even the evaluation order differs from the two other oracles.
tests/test_autograd_oracle.py compares it with ude_oracle.c on every loss x right-hand side x solver of tests/cases.py.
"""
from __future__ import annotations

import numpy as np
import torch

_TABLEAUX = {
    0: ([0.0, 0.5, 0.5, 1.0], [[], [0.5], [0.0, 0.5], [0.0, 0.0, 1.0]], [1 / 6, 1 / 3, 1 / 3, 1 / 6]),
    1: ([0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0],
        [[], [0.161], [-0.008480655492356989, 0.335480655492357],
         [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
         [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
         [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]],
        [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]),
    2: ([0.0], [[]], [1.0]),
}


def _interp(p, series, t):
    """covariates(t, series): per-variable piecewise-linear, flat outside the knots (src/covariates.jl:16-19, :45-48).
    series, t: 1-D numpy arrays of equal length -> (len, n_cov) float64."""
    out = np.zeros((len(t), p.n_cov))
    for i, (s, ti) in enumerate(zip(series, t)):
        for v in range(p.n_cov):
            k0, k1 = int(p.knot_off[s * p.n_cov + v]), int(p.knot_off[s * p.n_cov + v + 1])
            out[i, v] = np.interp(ti, p.knot_t[k0:k1], p.knot_x[k0:k1]) if k1 > k0 else 0.0
    return out


class _Model:
    """process model of ONE theta block, evaluated on a batch of states (rows)."""

    def __init__(self, p, pm, series_param=None):
        H, din, dout = p.nn_hidden, p.nn_in, p.nn_out
        self.p = p
        self.W1 = pm[p.off_w1:p.off_w1 + H * din].reshape(din, H).T          # column-major (H x din)
        self.b1 = pm[p.off_b1:p.off_b1 + H]
        self.W2 = pm[p.off_w2:p.off_w2 + dout * H].reshape(H, dout).T        # column-major (dout x H)
        self.b2 = pm[p.off_b2:p.off_b2 + dout]
        self.sc = [pm[o] for o in p.off_scalar]
        self.sp = pm[p.off_series_param:] if p.has_series_param else None

    def nn(self, x):
        return torch.tanh(x @ self.W1.T + self.b1) @ self.W2.T + self.b2

    def rhs(self, u, t, series_g, series_l):
        """u: (B, d) tensor; t: (B,) numpy; series_*: (B,) numpy int -> du (B, d)"""
        p, k = self.p, self.p.rhs_kind
        X = torch.from_numpy(_interp(p, series_g, t)) if p.n_cov else None
        if k in (0, 4):
            x = torch.cat([u, X], dim=1) if p.n_cov else u
        elif k == 1:
            tf = torch.from_numpy(p.time_scale * t + p.time_shift).reshape(-1, 1)
            x = torch.cat([u] + ([X] if p.n_cov else []) + [tf], dim=1)
        else:
            x = u
        y = self.nn(x)
        sc = self.sc
        if k == 0:
            return y
        if k == 1:
            dout = p.nn_out
            return torch.cat([y] + [(-sc[j - dout] * u[:, j]).reshape(-1, 1) for j in range(dout, p.d)], dim=1)
        if k == 2:
            du0 = sc[0] * u[:, 0] - y[:, 0]
            du1 = sc[2] * y[:, 0] - sc[1] * u[:, 1]
            if p.n_scalars >= 5:
                du0 = du0 + sc[3] * X[:, 0]
                du1 = du1 + sc[4] * X[:, 0]
            return torch.stack([du0, du1], dim=1)
        if k == 3:
            r, kk, m, th = (torch.abs(s) for s in sc[:4])
            C = torch.abs(y[:, 0])
            return torch.stack([r * u[:, 0] * (1.0 - u[:, 0] / kk) - C, th * C - m * u[:, 1]], dim=1)
        if k == 4:
            return torch.stack([y[:, 0], sc[1] * u[:, 1] * (1.0 - u[:, 1] / sc[2]) - sc[0] * u[:, 0]], dim=1)
        if k == 5:
            return y - sc[0] * u
        if k == 6:
            return self.sp[torch.from_numpy(np.asarray(series_l))].reshape(-1, 1) * u * y[:, :1]
        raise ValueError(k)

    def step(self, u, t, h, sg, sl, solver=None):
        c, a, b = _TABLEAUX[self.p.solver if solver is None else solver]
        hh = torch.from_numpy(np.asarray(h, dtype=np.float64)).reshape(-1, 1)
        ks = []
        for i in range(len(c)):
            U = u
            if i > 0:
                acc = sum(a[i][l] * ks[l] for l in range(i) if a[i][l] != 0.0)
                U = u + hh * acc
            ks.append(self.rhs(U, t + c[i] * h, sg, sl))
        return u + hh * sum(b[i] * ks[i] for i in range(len(c)))

    def flow(self, u, t0, t1, sg, sl):
        """one observation interval = `substeps` equal fixed steps (src/ProcessModels.jl:80-86 with a fixed discretisation)"""
        S = self.p.substeps
        h = (t1 - t0) / S
        for n in range(S):
            u = self.step(u, t0 + n * h, h, sg, sl)
        return u


def loss(p, theta: torch.Tensor) -> torch.Tensor:
    """loss[n_models] as a differentiable function of the flat theta (torch.float64)."""
    d = p.d
    out = []
    for m in range(p.n_models):
        base = int(p.theta_base[m])
        ncm = int(p.model_col0[m + 1] - p.model_col0[m])
        uh = theta[base:base + d * ncm].reshape(ncm, d)            # row t = column t of the d x N matrix
        pm = theta[base + d * ncm: base + d * ncm + p.n_pm]
        mdl = _Model(p, pm)
        s0, s1 = int(p.model_series0[m]), int(p.model_series0[m + 1])
        mc0 = int(p.model_col0[m])
        Y = torch.from_numpy(np.ascontiguousarray(p.data.T))         # (N, d)
        tt = p.times
        L = torch.zeros((), dtype=torch.float64)
        A = torch.from_numpy(np.ascontiguousarray(p.A)); B = torch.from_numpy(np.ascontiguousarray(p.B))
        if p.loss_kind == 0:
            # every observation term, then every (not skipped) interval of the model in ONE batched flow
            cols = np.arange(mc0, mc0 + ncm)
            e = Y[cols] - uh
            L = L + p.obs_scale[m] * torch.einsum("ti,ij,tj->", e, A, e)
            ends, sg = [], []
            for s in range(s0, s1):
                c0, n = int(p.starts[s]), int(p.lengths[s])
                for t in range(1, n):
                    if p.skip_mask is not None and p.skip_mask[c0 + t]:
                        continue
                    ends.append(c0 + t); sg.append(s)
            if ends:
                ends = np.array(ends); sg = np.array(sg)
                pred = mdl.flow(uh[ends - 1 - mc0], tt[ends - 1], tt[ends], sg, sg - s0)
                r = uh[ends - mc0] - pred
                w = p.proc_scale[m] * ((1.0 / (tt[ends] - tt[ends - 1])) if p.proc_inv_dt else np.ones(len(ends)))
                L = L + torch.einsum("t,ti,ij,tj->", torch.from_numpy(w), r, B, r)
        elif p.loss_kind == 1:
            # all blocks of the model advance together, interval by interval; a block that has run out of intervals is dropped
            blocks = []
            for s in range(s0, s1):
                c0, n = int(p.starts[s]), int(p.lengths[s])
                for t in range(0, n, p.pred_length):
                    Lb = p.pred_length if (n - 1 - t) >= p.pred_length else (n - 1 - t)
                    blocks.append((s, c0 + t, Lb, 1.0 / (d * n)))
            sg = np.array([b[0] for b in blocks]); c = np.array([b[1] for b in blocks]); Lb = np.array([b[2] for b in blocks])
            w = torch.from_numpy(np.array([b[3] for b in blocks]))
            u = uh[c - mc0]
            if p.preroll_eps > 0:
                u = mdl.step(u, tt[c] - p.preroll_eps, np.full(len(c), p.preroll_eps), sg, sg - s0)
            alive = np.arange(len(blocks))
            i = 0
            while len(alive):
                r = Y[c[alive] + i] - u
                L = L + torch.sum(w[alive] * torch.sum(r * r, dim=1))
                keep = Lb[alive] > i
                u = u[torch.from_numpy(keep)]
                alive = alive[keep]
                if len(alive):
                    u = mdl.flow(u, tt[c[alive] + i], tt[c[alive] + i + 1], sg[alive], sg[alive] - s0)
                i += 1
        elif p.loss_kind == 2:
            for s in range(s0, s1):
                c0, n = int(p.starts[s]), int(p.lengths[s])
                u = (uh[c0 - mc0] if p.shoot_from_theta else Y[c0]).reshape(1, d)
                sa = np.array([s])
                for i in range(n):
                    if i >= p.shoot_first_scored:
                        r = Y[c0 + i] - u[0]
                        L = L + p.shoot_weight[m] * torch.sum(r * r)
                    if i < n - 1:
                        u = mdl.flow(u, tt[c0 + i:c0 + i + 1], tt[c0 + i + 1:c0 + i + 2], sa, sa - s0)
        else:
            cols, sg = [], []
            for s in range(s0, s1):
                c0, n = int(p.starts[s]), int(p.lengths[s])
                for t in range(p.remove_ends, n - p.remove_ends):
                    cols.append(c0 + t); sg.append(s)
            if cols:
                cols = np.array(cols); sg = np.array(sg)
                U = torch.from_numpy(np.ascontiguousarray(p.dm_u.T))[cols]
                f = mdl.rhs(U, tt[cols], sg, sg - s0)
                r = torch.from_numpy(np.ascontiguousarray(p.dm_dudt.T))[cols] - f
                L = L + torch.sum(r * r)
        if p.reg_weight[m] != 0:
            L = L + p.reg_weight[m] * (torch.sum(mdl.W1 * mdl.W1) + torch.sum(mdl.W2 * mdl.W2))
        out.append(L)
    return torch.stack(out)


def loss_grad(p, theta):
    """(loss[n_models], grad[n_theta]) as numpy; the models are independent, so d(sum of losses)/d(theta) is the gradient."""
    th = torch.tensor(np.asarray(theta, dtype=np.float64), requires_grad=True)
    Ls = loss(p, th)
    Ls.sum().backward()
    return Ls.detach().numpy().copy(), th.grad.detach().numpy().copy()
