"""Host mirror of the reference's forward_solver.py for the simulation hot path.

``solve`` keeps the reference's signature and argument meaning (forward_solver.py:41-92) but is
a thin ctypes call into the CUDA library: NumPy arrays in, NumPy arrays out, no PyTorch and no
CPU fallback.  ``solve_batch`` is the same call for many parameter sets and measurements at once,
which is how the GPU wants to be fed.

Differences from the reference that a caller can observe (all deliberate):
  * solver=("solveivp",) / ("odeint",) both select the device integrator (RODAS4); ("NN", ...) is
    not part of this path and raises NotImplementedError.
  * ``state`` is not modified (the reference scales it in place and scales it back).
  * RTOL means the same thing and has the same default; ATOL above 1e-30 nm^-3 is clamped
    (``_capi.effective_tolerances``).  ``g.hmax`` is honoured only with ``honor_hmax=True``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _capi
from .sim_utils import Grid

DEFAULT_RTOL = _capi.DEFAULT_RTOL     # forward_solver.py:18
DEFAULT_ATOL = 1e-10                  # forward_solver.py:19 (accepted, clamped)
eps0 = 8.854 * 1e-12 * 1e-9           # forward_solver.py:21
q = 1.0
q_C = 1.602e-19                       # forward_solver.py:23
kB = 8.61773e-5                       # forward_solver.py:24

_CTX: dict = {}


def get_context(device: Optional[int] = None) -> _capi.Context:
    """Process-wide context per device (one process per GPU)."""
    import os
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("TRPL_USE_LOCAL_RANK") else 0
    if device not in _CTX:
        _CTX[device] = _capi.Context(device)
    return _CTX[device]


def E_field(N, P, n0, p0, eps, dx, corner_E=0):
    """Gauss's law on the host (forward_solver.py:26-38); the device does this with a warp scan."""
    N = np.asarray(N)
    if N.ndim not in (1, 2):
        raise NotImplementedError(f"Unsupported number of dimensions: {N.ndim}")
    E = corner_E + q_C / (eps * eps0) * dx * np.cumsum(((P - p0) - (N - n0)), axis=-1)
    pad = np.full(E.shape[:-1] + (1,), float(corner_E))
    return np.concatenate((pad, E), axis=-1)


def solve_batch(iniPars, grids, states, indexes, meas_types=None, units=None, model="std",
                ini_mode="density", RTOL=None, ATOL=None, honor_hmax=False, device=None,
                return_info=False):
    """Simulate every (state, measurement) pair.

    iniPars : sequence of per-measurement initial conditions (as in shared_fields["_init_params"])
    grids   : sequence of Grid, one per measurement
    states  : [n_sets, n_params] parameter sets in file units
    Returns a list (one entry per measurement) of [n_sets, len(grid.tSteps)] arrays.
    """
    n_meas = len(grids)
    if meas_types is None:
        meas_types = ["TRPL"] * n_meas
    sim_info = {"num_meas": n_meas, "lengths": [g.thickness for g in grids],
                "nx": [g.nx for g in grids], "meas_types": list(meas_types)}
    prob = _capi.pack_problem(sim_info, iniPars, [g.tSteps for g in grids], None, None, model=model,
                              ini_mode=ini_mode, min_y=[g.min_y for g in grids])
    params = _capi.pack_params(states, indexes, units, model=model)
    n_sets = params.shape[0]
    aux = _capi.default_aux(n_sets, n_meas, [1.0] * n_meas)
    hmax = min(g.hmax for g in grids)
    opts = _capi.make_opts(RTOL, ATOL, hmax=hmax, honor_hmax=honor_hmax,
                           flags=_capi.OPT_NO_LIKELIHOOD)
    ctx = get_context(device)
    ctx.set_problem(prob)
    curves, status, nsteps = ctx.solve_batch(params, aux, opts)
    out = [curves[:, prob.t_off[i]:prob.t_off[i] + prob.n_t[i]].copy() for i in range(n_meas)]
    if return_info:
        return out, status, nsteps
    return out


def solve(iniPar, g: Grid, state, indexes, meas="TRPL", units=None, solver=("solveivp",),
          model="std", ini_mode="density", RTOL=None, ATOL=None, honor_hmax=False):
    """One simulation, same call as the reference's solve() (forward_solver.py:41-203)."""
    if solver[0] == "NN":
        raise NotImplementedError("the NN surrogate is outside the CUDA hot path")
    if solver[0] not in ("solveivp", "odeint", "diagnostic"):
        raise NotImplementedError
    if meas not in ("TRPL", "TRTS"):
        raise NotImplementedError("TRTS or TRPL only")
    state = np.asarray(state, dtype=np.float64)
    out = solve_batch([np.asarray(iniPar, dtype=np.float64)], [g], state[None, :], indexes,
                      meas_types=[meas], units=units, model=model, ini_mode=ini_mode, RTOL=RTOL,
                      ATOL=ATOL, honor_hmax=honor_hmax)
    return out[0][0]
