"""The reference's README tutorial (README.md:40-75) on libude_cuda, line for line.

    julia                                                          here
    -----                                                          ----
    data, plt = UniversalDiffEq.LotkaVolterra()                    data = ude.LotkaVolterra()
    NN, params = UniversalDiffEq.SimpleNeuralNetwork(2, 1)          NN, params = ude.SimpleNeuralNetwork(2, 1)
    init_parameters = (NN = params, r = .1, k = 10.1, ...)         init_parameters = {"NN": params, "r": .1, "k": 10.1, ...}
    function dudt(u, p, t) ... end                                 dudt = ude.expression_rhs(...)   # the same formulas as expressions
    model = CustomDerivatives(data, dudt, init_parameters)          model = ude.CustomDerivatives(data, dudt, init_parameters)
    train!(model; loss_function = "derivative matching", ...)      ude.train_(model, loss_function="derivative matching", ...)

`dudt` is compiled into the fused RK + loss + adjoint kernels on first use; the hand-written catalog entry of the
same model (ude.LV_UDE_ABS) is trained beside it and must give the same parameters.
usage (on a B200): python examples/readme_tutorial.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ude_b200 as ude  # noqa: E402


def main():
    data = ude.LotkaVolterra()                                   # synthetic predator-prey data: time, x1, x2
    NN, params = ude.SimpleNeuralNetwork(2, 1)
    init_parameters = {"NN": params, "r": 0.1, "k": 10.1, "m": 0.1, "theta": 0.1}
    # C = abs(NN(u)[1]); r, k, theta, m = abs.(...); dx1 = r u1 (1 - u1 / k) - C; dx2 = theta C - m u2
    dudt = ude.expression_rhs(d=2, nn_out=1, scalars=("r", "k", "m", "theta"),
                              du=["abs(r)*u1*(1 - u1/abs(k)) - abs(y1)", "abs(theta)*abs(y1) - abs(m)*u2"])
    model = ude.CustomDerivatives(data, dudt, init_parameters)
    t0 = time.perf_counter()
    ude.train_(model, loss_function="derivative matching", optimizer="ADAM", regularization_weight=0.0, verbose=False,
               loss_options=dict(d=10), optim_options=dict(maxiter=1000, step_size=0.01))
    t_fit = time.perf_counter() - t0
    twin = ude.CustomDerivatives(data, ude.LV_UDE_ABS, init_parameters)          # the hand-written kernel of the same model
    ude.train_(twin, loss_function="derivative matching", optimizer="ADAM", regularization_weight=0.0, verbose=False,
               loss_options=dict(d=10), optim_options=dict(maxiter=1000, step_size=0.01))
    same = float(np.max(np.abs(model.parameters - twin.parameters)) / np.max(np.abs(twin.parameters)))
    # second pass of the tutorial: conditional likelihood (state-space) fit, then one-step-ahead predictions on the last 10 points
    t0 = time.perf_counter()
    ude.train_(model, loss_function="conditional likelihood", verbose=False, optim_options=dict(maxiter=500, step_size=0.05))
    t_fit2 = time.perf_counter() - t0
    inits, obs, preds = ude.predictions(model, data.iloc[-10:])
    rmse = float(np.sqrt(np.mean((preds - obs) ** 2)))
    print(f"derivative matching, 1000 Adam iterations: {t_fit:.2f} s (incl. generating and compiling dudt on first use)")
    print(f"generated kernels vs hand-written catalog kernel after 1000 iterations: max relative parameter difference {same:.1e}")
    print(f"conditional likelihood, 500 Adam iterations: {t_fit2:.2f} s; one-step-ahead RMSE on the last 10 points: {rmse:.3f}")
    pm = model.process_model
    print("fitted r, k, m, theta:", np.round(np.abs(pm[-4:]), 3))
    assert same < 1e-8 and np.isfinite(rmse)


if __name__ == "__main__":
    main()
