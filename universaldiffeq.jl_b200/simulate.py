"""Simulated data sets with the reference's names and signatures (/root/reference/src/SimulateData.jl:21-55, :80-118,
:244-276), returned as pandas DataFrames with the same columns.  The noise comes from NumPy's default_rng(seed): Julia's
Random.seed! streams cannot be reproduced, so the draws differ from the reference's while the distributions agree.
`plot` is accepted and ignored (plotting is out of scope)."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import synthetic


def LotkaVolterra(plot=False, seed=123, datasize=60, T=3.0, sigma=0.075) -> pd.DataFrame:
    """du1 = r u1 - alpha u1 u2, du2 = theta alpha u1 u2 - m u2 with (4, 5, .75, 5), u0 = (1, .5), multiplicative N(0, sigma)
    observation noise; columns time, x1, x2."""
    rng = np.random.default_rng(seed)
    Y, ts = synthetic.lotka_volterra_data(rng, n_series=1, n_pts=datasize, T=T, sigma=sigma)
    return pd.DataFrame({"time": ts, "x1": Y[0, 0], "x2": Y[0, 1]})


def LorenzLotkaVolterra(plot=False, seed=123, datasize=60, T=3.0, sigma=0.075):
    """Logistic predator-prey pair forced by the x-coordinate of a Lorenz system; returns (data[time,x1,x2], X[time,X])."""
    rng = np.random.default_rng(seed)
    ts = np.linspace(0.0, T, datasize)
    r, K, alpha, theta, m, sg, rho, beta = 4.0, 10.0, 5.0, 0.75, 5.0, 10.0, 28.0, 8.0 / 3.0

    def f(u, t):
        du = np.empty_like(u)
        du[:, 0] = r * u[:, 0] * (1 - u[:, 0] / K) - alpha * u[:, 0] * u[:, 1] + 0.05 * u[:, 2]
        du[:, 1] = theta * alpha * u[:, 0] * u[:, 1] - m * u[:, 1]
        du[:, 2] = sg * (u[:, 3] - u[:, 2])
        du[:, 3] = u[:, 2] * (rho - u[:, 4]) - u[:, 2]        # as written in the reference (not the textbook -u4)
        du[:, 4] = u[:, 2] * u[:, 3] - beta * u[:, 4]
        return du

    clean = synthetic._rk4_integrate(f, np.array([[1.0, 0.5, -10.0, -5.0, 20.0]]), ts, sub=200)[0]
    noisy = clean + clean * rng.normal(0.0, sigma, size=clean.shape)
    return (pd.DataFrame({"time": ts, "x1": noisy[0], "x2": noisy[1]}), pd.DataFrame({"time": ts, "X": noisy[2]}))


def simulate_coral_data():
    """Macroalgae / coral cover with an autocorrelated environmental driver; returns (data[time,A,C], X[time,X]),
    125 points at time = 1, 3, ..., 249."""
    rng = np.random.default_rng(1)
    Y, ts = synthetic.coral_data(rng, n_pts=125)
    return pd.DataFrame({"time": ts, "A": Y[0], "C": Y[1]}), pd.DataFrame({"time": ts, "X": Y[2]})
