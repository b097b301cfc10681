"""Host side of the IRF convolution: the one-off moment tables.

make_I_tables mirrors laplace.py:13-41 / I_moment laplace.py:132-175 of the reference (same SciPy
Simpson rule on the same 1000-point interpolants, so the tables are bit-identical); the per-curve
work - resampling, convolution, max-shift, trim (laplace.py:44-129, :178-222) - runs inside the
CUDA kernel (csrc/irf.h).
"""
import os

import numpy as np
from scipy.integrate import simpson


def I_moment(t, y, m, n, u_lower=0, u_upper=1, u_spacing=100):
    """Moment integral I_m^n of an IRF sampled on the regular grid t."""
    dt = t[1] - t[0]
    u = np.linspace(u_lower, u_upper, u_spacing)
    du = u[1] - u[0]
    y_lin = np.linspace(y[m + 1 - u_lower], y[m + 1 - u_upper], u_spacing)
    return dt * simpson((u - 0.5) ** n * y_lin, dx=du)


def make_I_tables(irfs):
    """{wavelength: raw (t, IRF(t)) array} -> {wavelength: (moments[nk, 3], t_irf)}."""
    tables = {}
    for w, irf in irfs.items():
        w = int(w)
        t_irf = irf[:, 0]
        f_irf = irf[:, 1]
        nk = len(f_irf)
        mom = np.zeros((nk, 3))
        for m in range(nk - 1):
            for n in range(3):
                mom[m, n] = I_moment(t_irf, f_irf, m, n, u_spacing=1000)
        tables[w] = (mom, t_irf)
    return tables


def load_irf_tables(wavelengths, irf_dir="IRFs"):
    """metropolis.py:331-338: read IRFs/irf_<w>nm.csv for every wavelength in use."""
    irfs = {}
    for w in wavelengths:
        if w > 0 and int(w) not in irfs:
            irfs[int(w)] = np.loadtxt(os.path.join(irf_dir, f"irf_{int(w)}nm.csv"), delimiter=",")
    return make_I_tables(irfs)
