"""Randomised differential test of the UKF marginal likelihood (ude_ukf_loss_grad / ude_ukf_filter, single- and many-model
handles, both prediction windows) against the NumPy twin with complex-step gradients.
usage (under gpurun): python tools/fuzz_ukf.py [--n 120] [--seed 0]"""
import argparse
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=120)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import ude_b200 as ude
    import cases
    from oracle import ukf_oracle_np as ukf
    rng = np.random.default_rng(a.seed)
    names = [n for n, v in cases.RHS_SPECS.items() if not v[6] and v[1] <= 4]
    fails, refused = 0, 0
    for it in range(a.n):
        rhs = names[int(rng.integers(len(names)))]
        d = cases.RHS_SPECS[rhs][1]
        solver = int(rng.integers(2))
        hidden = int(rng.choice([2, 3, 5, 8, 10, 12, 17]))
        M = int(rng.choice([1, 1, 2, 3]))
        per = int(rng.integers(1, 4))
        lengths = []
        for m in range(M):
            ls = [int(x) for x in rng.choice([1, 2, 3, 4, 6, 8], size=per)]
            if max(ls) < 2:
                ls[0] = int(rng.choice([2, 3, 5]))
            lengths += ls
        causal = bool(rng.integers(2))
        alpha = float(rng.choice([1.0, 1.0, 0.5]))
        kw = dict(hidden=hidden, lengths=tuple(lengths), n_models=M, seed=int(rng.integers(1 << 30)), substeps=int(rng.choice([1, 2, 4])))
        try:
            prob, theta = cases.make_case(rhs, "cond_lik", solver, **kw)
            Fs = np.stack([0.3 * np.eye(d) + 0.05 * rng.normal(size=(d, d)) for _ in range(M)])
            Peta = 0.1 * np.eye(d) + 0.02 * np.ones((d, d))
            with ude.Engine(prob) as eng:
                L, g, gF = eng.ukf_loss_grad(theta, Fs if M > 1 else Fs[0], Peta, alpha=alpha, reg=2e-3, causal=causal)
                xs, Ps = eng.ukf_filter(theta, Fs if M > 1 else Fs[0], Peta, alpha=alpha, causal=causal)
            L = np.atleast_1d(L); gF = np.asarray(gF).reshape(M, d, d)
            worst = 0.0
            for m in range(M):
                mm = m if M > 1 else None
                ref = ukf.loss(prob, theta, Fs[m], Peta, alpha=alpha, reg_weight=1e-3, reg_mult=2.0, causal=causal, model=mm)
                g_pm, g_F = ukf.complex_step_grad(prob, theta, Fs[m], Peta, alpha=alpha, reg_weight=1e-3, reg_mult=2.0, causal=causal, model=mm)
                off = ukf._model_range(prob, mm)[0]
                worst = max(worst, abs(L[m] - ref) / abs(ref), cases.relerr(g[off:off + prob.n_pm], g_pm), cases.relerr(gF[m].reshape(-1, order="F"), g_F))
                s0, s1 = ukf._model_range(prob, mm)[1:]
                for s in range(s0, s1):
                    _, xr, _ = ukf.filter_series(prob, theta[off:off + prob.n_pm], s, Fs[m] @ Fs[m].T, Peta, alpha, 2.0, 0.0, causal=causal)
                    c0, n = int(prob.starts[s]), int(prob.lengths[s])
                    worst = max(worst, cases.relerr(xs[:, c0:c0 + n], xr))
            tol = 1e-10 if alpha == 1.0 else 4e-10
            if not (worst < tol):
                fails += 1
                print("FAIL", it, rhs, solver, kw, "causal", causal, "alpha", alpha, "worst", worst, flush=True)
        except ude.UdeError as ex:
            if ex.code == -2:                      # UDE_JIT=0 and no shipped kernel for the shape: a documented refusal
                refused += 1
                continue
            fails += 1
            print("ERROR", it, rhs, solver, kw, repr(ex), flush=True)
        except Exception:
            fails += 1
            print("EXC", it, rhs, solver, kw, flush=True)
            traceback.print_exc()
    print("ukf cases", a.n, "refused", refused, "failures", fails)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
