"""Maximum-likelihood driver, mirroring MLE/max_likelihood.py.

The reference minimises the negative log-likelihood of one state with SciPy's Nelder-Mead, one
forward simulation of every measurement per function call (max_likelihood.py:11-110, :148-150).
Nelder-Mead is sequential, but what it may ask for next is known in advance: the reflection, the
expansion and the two contractions of the current simplex (or, after a failed contraction, the
N shrunk vertices).  ``nelder_mead_batched`` evaluates those candidates in ONE batched likelihood
call per iteration and then takes exactly the decisions SciPy's ``_minimize_neldermead`` takes
(same coefficients, initial simplex, ordering and termination tests), so the iterates are SciPy's
as long as the function values are.  ``mle`` keeps the reference's signature.
"""
from __future__ import annotations

import os

import numpy as np

from .laplace import load_irf_tables
from .sim_utils import Ensemble

DEFAULT_NUM_ITERS = 1000        # max_likelihood.py:10


def nelder_mead_batched(fun_batch, x0, xatol=1e-4, fatol=1e-4, maxiter=None, maxfev=None, callback=None):
    """Nelder-Mead with SciPy's rules; ``fun_batch(X[n, N]) -> f[n]`` evaluates candidates together.

    Returns a dict with x, fun, nit, nfev (the evaluations SciPy would have made), nbatch (the
    batched calls actually made), npoints (points actually evaluated) and message.
    """
    rho, chi, psi, sigma = 1.0, 2.0, 0.5, 0.5
    x0 = np.asarray(x0, dtype=float).ravel()
    N = len(x0)
    sim = np.empty((N + 1, N))
    sim[0] = x0
    for k in range(N):                       # scipy: nonzdelt = 0.05, zdelt = 0.00025
        y = x0.copy()
        y[k] = (1 + 0.05) * y[k] if y[k] != 0 else 0.00025
        sim[k + 1] = y
    if maxiter is None and maxfev is None:
        maxiter = maxfev = N * 200
    elif maxiter is None:
        maxiter = N * 200 if maxfev == np.inf else np.inf
    elif maxfev is None:
        maxfev = N * 200 if maxiter == np.inf else np.inf
    nbatch, npoints = 1, N + 1
    fsim = np.asarray(fun_batch(sim), dtype=float)
    nfev = N + 1
    ind = np.argsort(fsim)
    sim, fsim = sim[ind], fsim[ind]
    nit = 1
    while nfev < maxfev and nit < maxiter:
        if np.max(np.ravel(np.abs(sim[1:] - sim[0]))) <= xatol and np.max(np.abs(fsim[0] - fsim[1:])) <= fatol:
            break
        xbar = np.add.reduce(sim[:-1], 0) / N
        xr = (1 + rho) * xbar - rho * sim[-1]
        xe = (1 + rho * chi) * xbar - rho * chi * sim[-1]
        xc = (1 + psi * rho) * xbar - psi * rho * sim[-1]
        xcc = (1 - psi) * xbar + psi * sim[-1]
        fxr, fxe, fxc, fxcc = np.asarray(fun_batch(np.stack([xr, xe, xc, xcc])), dtype=float)
        nbatch += 1
        npoints += 4
        nfev += 1
        doshrink = False
        if fxr < fsim[0]:
            nfev += 1
            if fxe < fxr:
                sim[-1], fsim[-1] = xe, fxe
            else:
                sim[-1], fsim[-1] = xr, fxr
        elif fxr < fsim[-2]:
            sim[-1], fsim[-1] = xr, fxr
        elif fxr < fsim[-1]:
            nfev += 1
            if fxc <= fxr:
                sim[-1], fsim[-1] = xc, fxc
            else:
                doshrink = True
        else:
            nfev += 1
            if fxcc < fsim[-1]:
                sim[-1], fsim[-1] = xcc, fxcc
            else:
                doshrink = True
        if doshrink:
            sim[1:] = sim[0] + sigma * (sim[1:] - sim[0])
            fsim[1:] = np.asarray(fun_batch(sim[1:]), dtype=float)
            nbatch += 1
            npoints += N
            nfev += N
        ind = np.argsort(fsim)
        sim, fsim = sim[ind], fsim[ind]
        if callback is not None:
            callback(sim[0], fsim[0])
        nit += 1
    if nfev >= maxfev:
        message = "Maximum number of function evaluations has been exceeded."
    elif nit >= maxiter:
        message = "Maximum number of iterations has been exceeded."
    else:
        message = "Optimization terminated successfully."
    return {"x": sim[0], "fun": float(fsim[0]), "nit": nit, "nfev": nfev, "nbatch": nbatch,
            "npoints": npoints, "message": message, "final_simplex": (sim, fsim)}


def mle(e_data, sim_params, param_info, init_params, sim_flags, export_path, logger, evaluator=None,
        irf_dir="IRFs", device=None, kernel="warp"):
    """Same call and result as max_likelihood.py:113-160: the Ensemble whose single chain holds the
    visited states (column k = k-th accepted best vertex) and their log-likelihoods.

    evaluator(states[n, n_params]) -> logll[n] may be injected (tests); by default the CUDA path.
    kernel="seulex": the low-latency integrator (PathCache) - a simplex iteration is a handful of
    trajectories, i.e. pure latency.
    """
    names = list(param_info["names"])
    sigma = sim_flags.get("current_sigma", None) or sim_flags.get("model_uncertainty", None)
    if sigma is None:
        raise KeyError("current_sigma")            # max_likelihood.py:95 reads MCMC_fields["current_sigma"]
    sim_flags = dict(sim_flags)
    sim_flags.setdefault("num_iters", DEFAULT_NUM_ITERS)
    sim_flags.setdefault("checkpoint_freq", DEFAULT_NUM_ITERS)
    MS_list = Ensemble(param_info, sim_params, sim_flags, DEFAULT_NUM_ITERS)
    ef = MS_list.ensemble_fields
    ef["_init_params"] = init_params
    ef["_times"], ef["_vals"], ef["_uncs"] = e_data
    if ef.get("irf_convolution", None) is not None:
        ef["_IRF_tables"] = load_irf_tables(ef["irf_convolution"], irf_dir)
    else:
        ef["_IRF_tables"] = None
    if not ef["log_y"]:
        raise NotImplementedError("the likelihood kernel compares log10 signals (log_y = 1)")
    active = ef["active"]
    base = MS_list.H.states[0, :, 0].copy()
    if evaluator is None:
        from .trial_move_evaluation import PathCache, eval_trial_moves
        cache = PathCache(ef, device=device, kernel=kernel)

        def evaluator(states):
            return eval_trial_moves(states, np.ones(len(states)), sigma, ef, cache=cache).logll

    def cost_batch(X):
        states = np.repeat(base[None, :], len(X), axis=0)
        states[:, active] = 10 ** np.asarray(X)
        ll = np.asarray(evaluator(states), dtype=float)
        return np.where(np.isfinite(ll), -ll, np.inf)

    H = MS_list.H
    MS_list.latest_iter = 1

    def record(x, f):
        k = MS_list.latest_iter
        if k >= H.accept.shape[1]:
            H.extend(2 * H.accept.shape[1])
        H.states[0, :, k] = H.states[0, :, k - 1]
        H.states[0, active, k] = 10 ** x
        H.loglikelihood[0, k] = -f
        if logger is not None:
            logger.info(f"Iter {k} Cost: {f}")
        MS_list.latest_iter = k + 1

    x0 = np.log10(base[active])
    opt = nelder_mead_batched(cost_batch, x0, callback=record)
    if logger is not None:
        logger.info(10 ** opt["x"])
        logger.info(-opt["fun"])
        logger.info(opt["message"])
    MS_list.opt = {k: v for k, v in opt.items() if k != "final_simplex"}
    if MS_list.latest_iter < H.accept.shape[1]:
        H.truncate(MS_list.latest_iter)
    if export_path is not None:
        os.makedirs(ef["output_path"], exist_ok=True)
        MS_list.checkpoint(os.path.join(ef["output_path"], export_path))
    return MS_list
