"""The reference-facing host API at BASELINE C3 scale: MultiNODE on 10^5 series with a covariate, constructor + 20 device-resident
Adam iterations of the multiple-shooting fit + the default (derivative matching) fit, wall-clock per phase."""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ude_b200 as ude  # noqa: E402

n, T = 100_000, 30
rng = np.random.default_rng(0)
Y, ts = ude.synthetic.lotka_volterra_data(rng, n, T, random_u0=True, pairs=2)
kt, kx = ude.synthetic.ou_covariates(rng, n, T, ts[-1])
out = {}
t0 = time.perf_counter()
df = pd.DataFrame({"time": np.tile(ts, n), "series": np.repeat(np.arange(n), T)})
for j in range(4):
    df[f"x{j + 1}"] = Y[:, j, :].reshape(-1)
X = pd.DataFrame({"time": np.tile(kt, n), "series": np.repeat(np.arange(n), T), "X": kx.reshape(-1)})
out["frames_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); m = ude.MultiNODE(df, X, hidden_units=32, seed=1); out["constructor_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
ude.train_(m, "multiple shooting", verbose=False, optim_options=dict(maxiter=20))
out["train_multiple_shooting_20_iterations_s"] = time.perf_counter() - t0
out["loss_first_last"] = [float(m.extra["loss_trace"].ravel()[0]), float(m.extra["loss_trace"].ravel()[-1])]
t0 = time.perf_counter()
ude.train_(m, verbose=False, optim_options=dict(maxiter=20))          # default loss of train!(::MultiUDE): derivative matching
out["train_derivative_matching_20_iterations_s"] = time.perf_counter() - t0
print(json.dumps(out, indent=1))
