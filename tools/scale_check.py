"""64-bit indexing / scale sanity: the C3 problem tiled K times (K x 10^5 series, K x 3*10^6 columns, K x 0.96 GB of states) must
give K x the loss and K x the process-model gradient of the single copy, and the same uhat gradient in every copy.
usage (under gpurun): python tools/scale_check.py [--k 10]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--series", type=int, default=100_000)
    a = ap.parse_args()
    import ude_b200 as ude
    P = ude.problem
    p1, th1 = ude.synthetic.config_c3(n_series=a.series)
    with ude.Engine(p1) as eng:
        L1, g1 = eng.loss_grad(th1)
    K = a.k
    pk = P.Problem(d=p1.d, n_cov=p1.n_cov, nn_in=p1.nn_in, nn_hidden=p1.nn_hidden, nn_out=p1.nn_out, rhs_kind=p1.rhs_kind, solver=p1.solver,
                   substeps=p1.substeps, loss_kind=p1.loss_kind, pred_length=p1.pred_length, preroll_eps=p1.preroll_eps).default_layout()
    n1 = p1.n_cols
    pk.data = np.asfortranarray(np.tile(np.asarray(p1.data), (1, K)))
    pk.times = np.tile(np.asarray(p1.times), K)
    pk.starts = np.concatenate([np.asarray(p1.starts) + j * n1 for j in range(K)])
    pk.lengths = np.tile(np.asarray(p1.lengths), K)
    ko = np.asarray(p1.knot_off)
    pk.knot_off = np.concatenate([ko[:-1] + j * ko[-1] for j in range(K)] + [[K * ko[-1]]])
    pk.knot_t = np.tile(np.asarray(p1.knot_t), K); pk.knot_x = np.tile(np.asarray(p1.knot_x), K)
    pk.reg_weight = np.asarray(p1.reg_weight) * 1.0
    pk.finalize()
    d = p1.d
    thk = np.concatenate([np.tile(th1[:d * n1], K), th1[d * n1:]])
    t0 = time.perf_counter()
    with ude.Engine(pk) as eng:
        Lk, gk = eng.loss_grad(thk)
        steps = eng.steps_per_eval()
    dt = time.perf_counter() - t0
    reg1 = float(p1.reg_weight[0]) if np.ndim(p1.reg_weight) else float(p1.reg_weight)
    pm = th1[d * n1:]
    w = np.concatenate([pm[p1.off_w1:p1.off_w1 + p1.nn_hidden * p1.nn_in], pm[p1.off_w2:p1.off_w2 + p1.nn_out * p1.nn_hidden]])
    reg_term = reg1 * float(w @ w)                         # the regulariser enters once, not K times
    eL = abs(Lk[0] - (K * (L1[0] - reg_term) + reg_term)) / abs(Lk[0])
    gu = gk[:d * n1 * K].reshape(K, d * n1)
    e_u = float(np.max(np.abs(gu - g1[None, :d * n1])) / np.max(np.abs(g1[:d * n1])))
    gpm1 = g1[d * n1:].copy(); gpmk = gk[d * n1 * K:].copy()
    greg = np.zeros_like(gpm1)
    greg[p1.off_w1:p1.off_w1 + p1.nn_hidden * p1.nn_in] = 2 * reg1 * pm[p1.off_w1:p1.off_w1 + p1.nn_hidden * p1.nn_in]
    greg[p1.off_w2:p1.off_w2 + p1.nn_out * p1.nn_hidden] = 2 * reg1 * pm[p1.off_w2:p1.off_w2 + p1.nn_out * p1.nn_hidden]
    e_pm = float(np.max(np.abs(gpmk - (K * (gpm1 - greg) + greg))) / np.max(np.abs(gpmk)))
    out = {"k": K, "series": K * a.series, "columns": int(pk.n_cols), "theta_doubles": int(pk.n_theta), "state_bytes": int(8 * d * pk.n_cols),
           "rk_steps": int(steps), "host_call_s": dt, "loss_rel_err_vs_k_times": eL, "uhat_grad_rel_err_between_copies": e_u,
           "pm_grad_rel_err_vs_k_times": e_pm, "ok": bool(eL < 1e-11 and e_u < 1e-12 and e_pm < 1e-10)}
    print(json.dumps(out))
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
