"""Diffusion-coefficient estimators applied to the (time, MSD) rows the ``msd`` analysis
returns.  Host-side post-processing, as in pybilt/diffusion/diffusion_coefficients.py:17-167
(Einstein relation, linear fit, anomalous-diffusion fit); unlike the reference,
``anomalous_fit`` does not modify the ``times`` array it is given.
"""
from __future__ import annotations

import numpy as np
from scipy import optimize, stats


def diffusion_coefficient_Einstein(times, msd_vals, dim=2, time_range=None):
    t, y = _window(times, msd_vals, time_range)
    dt = t[-1] - t[0]
    return y[-1] / (2.0 * dim * dt)


def diffusion_coefficient_linear_fit(times, msd_vals, dim=2, time_range=None):
    t, y = _window(times, msd_vals, time_range)
    slope, _intercept, _r, _p, std_err = stats.linregress(t - t[0], y)
    return slope / (2.0 * dim), std_err / (2.0 * dim)


def diffusion_coefficient_anomalous_fit(times, msd_vals, dim=2, time_range=None):
    """Fit MSD = 2*dim*D_alpha*t**alpha; returns array (D_alpha, alpha)."""
    t, y = _window(times, msd_vals, time_range)
    t = t - t[0]
    prefactor = 2.0 * dim

    def model(time, d_alpha, alpha):
        return prefactor * d_alpha * time ** alpha

    params, _cov = optimize.curve_fit(model, t, y)
    return params


def _window(times, vals, time_range):
    t = np.asarray(times, dtype=np.float64)
    y = np.asarray(vals, dtype=np.float64)
    if time_range is not None:
        keep = (t >= time_range[0]) & (t <= time_range[1])
        t, y = t[keep], y[keep]
    return t, y
