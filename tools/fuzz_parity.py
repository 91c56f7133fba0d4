"""Randomised differential test: CUDA (through the C ABI) against the C oracle on random small problems -- right-hand side,
loss form, solver, hidden width (so: kernel family and geometry), substeps, ragged series lengths, models per handle, skip
masks, prediction-block length.  usage (under gpurun): python tools/fuzz_parity.py [--n 300] [--seed 0]
Prints one line per failure and a summary; exit code 1 if anything failed."""
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
    ap.add_argument("--n", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--dtype", default="f64", help="f64 (bar 1e-10, per gradient leaf) or f32 (the optional FP32 mode: bar 1e-4, norm-wise)")
    ap.add_argument("--forward", action="store_true", help="also compare ude_forward trajectories with the oracle's")
    ap.add_argument("--sparse", action="store_true", help="also run the zero-copy host path (UDE_OPT_SPARSE_UHAT_IO, page-locked buffers): the bench's e2e path")
    ap.add_argument("--adam", action="store_true", help="also compare 5 device-resident Adam iterations (ude_adam) with a host loop on the oracle's gradients")
    ap.add_argument("--prebuild", action="store_true", help="no GPU: only compile the on-demand kernel plugins the same random cases may need")
    a = ap.parse_args()
    import ude_b200 as ude
    import cases
    from oracle import oracle
    rng = np.random.default_rng(a.seed)
    names = list(cases.RHS_SPECS)
    fails, kernels, unsup = 0, {}, set()
    for it in range(a.n):
        rhs = names[int(rng.integers(len(names)))]
        loss = cases.LOSSES[int(rng.integers(len(cases.LOSSES)))]
        solver = int(rng.integers(2))
        d = cases.RHS_SPECS[rhs][1]
        hmax = 128 if rhs in ("node_d8", "node_d4x1", "node_d4", "node_d2") else 32
        hidden = int(rng.choice([1, 2, 3, 5, 7, 8, 10, 12, 16, 17, 24, 31, 32] + ([33, 64, 100, 128] if hmax == 128 else [])))
        n_models = int(rng.choice([1, 1, 1, 2, 3]))
        n_series = n_models * int(rng.integers(1, 5))
        lengths = tuple(int(x) for x in rng.choice([1, 2, 3, 4, 5, 6, 9, 11, 17, 30], size=n_series))
        substeps = int(rng.choice([1, 2, 3, 4]))
        pl = int(rng.choice([1, 2, 5, 7]))
        kw = dict(hidden=hidden, lengths=lengths, n_models=n_models, seed=int(rng.integers(1 << 30)), substeps=substeps,
                  skip=bool(rng.integers(2)) and loss in ("cond_lik", "mse"), pred_length=pl)
        if cases.RHS_SPECS[rhs][6] and n_models > 1 and n_series % n_models:
            continue
        if a.prebuild:
            if hidden >= 13:
                prob, theta = cases.make_case(rhs, loss, solver, **kw)
                ude.user_rhs.ensure_catalog_kernels(prob, 1 if a.dtype == "f32" else 0)
            continue
        try:
            prob, theta = cases.make_case(rhs, loss, solver, **kw)
            L_ref, g_ref = oracle.loss_grad(prob, theta)
            if not (np.all(np.isfinite(L_ref)) and np.all(np.isfinite(g_ref))):
                kernels["diverged in the oracle too (open-loop trajectory blows up): skipped"] = kernels.get("diverged in the oracle too (open-loop trajectory blows up): skipped", 0) + 1
                continue
            f32 = a.dtype == "f32"
            tol = 1e-4 if f32 else 1e-10
            fwd_err = 0.0
            with ude.Engine(prob, dtype=1 if f32 else 0) as eng:
                kn = eng.kernel_name()
                L, g = eng.loss_grad(theta)
                ok_steps = eng.steps_per_eval() == oracle.count_steps(prob)
                if a.forward and loss != "deriv_match":
                    out = eng.forward(theta)
                    ref = oracle.forward(prob, theta)
                    msk = np.isfinite(ref)
                    fwd_err = float(np.max(np.abs(out[msk] - ref[msk])) / max(float(np.max(np.abs(ref[msk]))), 1e-300)) if msk.any() else 0.0
                extra_ok = True
                if a.sparse and n_models == 1 and not f32:
                    import torch
                    eng.set_option(ude.OPT_SPARSE_UHAT_IO, 1)
                    h_theta = torch.from_numpy(theta.copy()).pin_memory()
                    h_grad = torch.zeros(prob.n_theta, dtype=torch.float64).pin_memory()
                    h_loss = torch.zeros(1, dtype=torch.float64).pin_memory()
                    for rep in range(2):
                        eng.loss_grad_ptr(h_theta.data_ptr(), prob.n_theta, h_loss.data_ptr(), h_grad.data_ptr())
                        gs = h_grad.numpy()
                        extra_ok = extra_ok and abs(h_loss[0].item() - L_ref[0]) < tol * abs(L_ref[0]) \
                            and cases.grad_mismatch(prob, gs, g_ref, rtol=tol)[0] <= 1.0 and bool(np.all(gs[g_ref == 0.0] == 0.0))
                    eng.set_option(ude.OPT_SPARSE_UHAT_IO, 0)
                if a.adam and not f32:
                    th_d, tr_d = eng.adam(theta, 1e-3, 5)
                    th_h = theta.copy(); mm = np.zeros_like(th_h); vv = np.zeros_like(th_h); tr_h = []
                    for k in range(1, 6):
                        Lh, gh = oracle.loss_grad(prob, th_h)
                        tr_h.append(Lh.copy())
                        mm = 0.9 * mm + 0.1 * gh; vv = 0.999 * vv + 0.001 * gh * gh
                        th_h -= 1e-3 * (mm / (1 - 0.9 ** k)) / (np.sqrt(vv / (1 - 0.999 ** k)) + 1e-8)
                    # Adam divides by sqrt(v): entries whose gradient is ~0 amplify round-off, so compare the traces tightly
                    # and the iterates at the step size's scale
                    extra_ok = extra_ok and float(np.max(np.abs(tr_d - np.array(tr_h)) / np.abs(np.array(tr_h)))) < 1e-9 \
                        and float(np.max(np.abs(th_d - th_h))) < 1e-3 * 5 * 1e-4
            kernels[kn] = kernels.get(kn, 0) + 1
            el = float(np.max(np.abs(L - L_ref) / np.abs(L_ref)))
            worst = (cases.relerr(g, g_ref) / tol, -1, "all") if f32 else cases.grad_mismatch(prob, g, g_ref, rtol=tol)
            zeros_ok = bool(np.all(g[g_ref == 0.0] == 0.0))
            if not (np.all(np.isfinite(g)) and el < tol and worst[0] <= 1.0 and zeros_ok and ok_steps and fwd_err < tol and extra_ok):
                fails += 1
                print("FAIL", it, rhs, loss, solver, kw, kn, "loss err", el, "grad", worst, "zeros", zeros_ok, "steps", ok_steps, "forward", fwd_err, "sparse/adam", extra_ok, flush=True)
        except ude.UdeError as ex:
            if ex.code == -2:                      # no compiled kernel for this shape: a documented refusal, not a failure
                kernels["unsupported"] = kernels.get("unsupported", 0) + 1
                unsup.add((rhs, hidden, n_models > 1))
                continue
            fails += 1
            print("ERROR", it, rhs, loss, solver, kw, repr(ex), flush=True)
        except Exception:
            fails += 1
            print("EXC", it, rhs, loss, solver, kw, flush=True)
            traceback.print_exc()
    print("cases", a.n, "failures", fails)
    for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]):
        print(f"  {v:4d}  {k}")
    print("unsupported (rhs, hidden, many-model):", sorted(unsup))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
