"""Dense (random-grid) sampling of the likelihood, mirroring Dense_Sample/dense_sampling.py.

The reference loops over measurements and samples one simulation at a time
(dense_sampling.py:105-196).  Here the sample grid is cut into blocks of parameter sets; each block
is one kernel launch over block x num_meas trajectories, and with several GPUs each rank takes a
contiguous slice of the grid (no collective until the final gather of the likelihood vector).
"""
from __future__ import annotations

import os

import numpy as np

from .laplace import load_irf_tables
from .parallel import Comm
from .utils import search_c_grps

DEFAULT_BLOCK = 16384     # parameter sets per launch: the tail of a launch is amortised (5% over 4096)

_SECOND_CTX = {}          # device -> the second context of the block pipeline (kept for the process)


def _second_context(device):
    """The pipeline's second context (stream + device buffers), created once per device: creating
    and destroying it on every call cost up to 0.4 s (cudaFree of its buffers) per simulate()."""
    from . import _capi
    if device not in _SECOND_CTX:
        _SECOND_CTX[device] = _capi.Context(device)
    return _SECOND_CTX[device]


def random_grid(min_X, max_X, do_log, num_samples):
    """Uniform (or log-uniform) draws inside the box, column by column (dense_sampling.py:17-35)."""
    n_par = len(min_X)
    grid = np.empty((num_samples, n_par))
    for i in range(n_par):
        if min_X[i] == max_X[i]:
            grid[:, i] = min_X[i]
        elif do_log[i]:
            grid[:, i] = 10 ** np.random.uniform(np.log10(min_X[i]), np.log10(max_X[i]), (num_samples,))
        else:
            grid[:, i] = np.random.uniform(min_X[i], max_X[i], (num_samples,))
    return grid


def make_grid(N, P, min_X, max_X, do_log, sim_flags):
    n = sim_flags["num_iters"]
    return np.arange(n), np.zeros(n), random_grid(min_X, max_X, do_log, n)


def modify_scale_factors(param_info, sim_flags):
    """dense_sampling.py:199-207."""
    spread = sim_flags["scale_factor"][0]
    for name in param_info["names"]:
        if name.startswith("_s"):
            g = param_info["init_guess"][name]
            param_info["prior_dist"][name] = (g / spread, g * spread)


def simulate(e_data, P, X, param_info, sim_params, init_params, sim_flags, logger=None,
             evaluator=None, comm=None, block=DEFAULT_BLOCK):
    """Fill P[i] with the log-likelihood of sample X[i] summed over measurements.

    Same result as dense_sampling.py:42-196 (solver 'solveivp', model 'std', hmax = 1 there is an
    LSODA safeguard and has no counterpart in the error-controlled integrator).
    """
    comm = comm or Comm()
    names = param_info["names"]
    units = np.array([param_info["unit_conversions"].get(p, 1) for p in names], dtype=float)
    idx = {name: names.index(name) for name in names}
    sf = {"_sim_info": sim_params, "_init_params": init_params, "_times": e_data[0], "_vals": e_data[1],
          "_uncs": e_data[2], "_param_indexes": idx, "units": units, "model": sim_flags.get("model", "std"),
          "ini_mode": sim_flags.get("ini_mode", "density"), "rtol": sim_flags.get("rtol", None),
          "atol": sim_flags.get("atol", None), "scale_factor": sim_flags.get("scale_factor", None),
          "irf_convolution": sim_flags.get("irf_convolution", None),
          "_IRF_tables": sim_flags.get("IRF_tables", None)}
    if not sim_flags["log_y"]:
        raise NotImplementedError("the likelihood kernel compares log10 signals (log_y = 1)")
    sigmas = sim_flags["current_sigma"]
    n = len(X)
    lo, hi = comm.shard(n)
    local = np.zeros(hi - lo)
    blocks = [(b0, min(b0 + block, hi)) for b0 in range(lo, hi, block)]
    if evaluator is None:
        # Two contexts (two streams, two sets of device buffers) take the blocks in turn: block
        # k+1 is uploaded and launched while block k still runs, so its CTAs fill the SMs that
        # block k's last trajectories leave idle and the copies hide behind the kernels.
        from . import _capi
        from .forward_solver import get_context
        from .trial_move_evaluation import PathCache
        dev = get_context(comm.local_rank).device
        caches = [PathCache(sf, device=dev), PathCache(sf, ctx=_second_context(dev))]
        pending = [None, None]

        def collect(k):
            b0, b1 = pending[k]
            per, _, _, _ = caches[k].ctx.download()
            ll = per[:, :, 0].sum(axis=1)
            local[b0 - lo:b1 - lo] = np.where(np.isnan(ll), -np.inf, ll)
            pending[k] = None

        for i, (b0, b1) in enumerate(blocks):
            k = i % 2
            if pending[k] is not None:
                collect(k)
            if logger is not None:
                logger.info(f"Rank {comm.rank}: samples {b0}..{b1} of {n}")
            c = caches[k]
            params, aux = c.pack(X[b0:b1], sigmas, np.ones((b1 - b0, 3)))
            c.ctx.set_problem_if_needed(c.prob)
            c.ctx.upload(params, aux)
            c.ctx.run_resident(c.opts())
            pending[k] = (b0, b1)
        for k in ((len(blocks)) % 2, (len(blocks) + 1) % 2):       # oldest first
            if pending[k] is not None:
                collect(k)
    else:
        for b0, b1 in blocks:
            if logger is not None:
                logger.info(f"Rank {comm.rank}: samples {b0}..{b1} of {n}")
            local[b0 - lo:b1 - lo] = evaluator(X[b0:b1])
    full = comm.allgather_rows(local[:, None], n)[:, 0] if comm.world > 1 else local
    P[:] += full
    return P


def bayes(N, P, init_params, sim_params, e_data, sim_flags, param_info, logger=None, evaluator=None,
          comm=None, irf_dir="IRFs"):
    """Driver with the reference's signature (dense_sampling.py:210-311)."""
    if sim_flags.get("scale_factor", None) is not None:
        modify_scale_factors(param_info, sim_flags)
    names = param_info["names"]
    act = param_info["active"]
    min_X = np.array([param_info["prior_dist"][n][0] if act[n] else param_info["init_guess"][n] for n in names])
    max_X = np.array([param_info["prior_dist"][n][1] if act[n] else param_info["init_guess"][n] for n in names])
    do_log = np.array([param_info["do_log"][n] for n in names])
    comm = comm or Comm()
    N, P, X = make_grid(N, P, min_X, max_X, do_log, sim_flags)
    # The grid comes from the unseeded global np.random (as in the reference, which never shards):
    # every rank would draw its own.  Rank 0's grid is the grid; the others receive it, so the
    # likelihood slices that simulate() gathers belong to the X that is returned and exported.
    X = comm.broadcast_array(X)
    if logger is not None:
        logger.info(f"Initializing {len(X)} random samples")
    param_info["trial_move"] = np.array([param_info["trial_move"][p] for p in names], dtype=float)
    l2m = sim_flags["likel2move_ratio"]
    top = max(param_info["trial_move"])
    sim_flags["current_sigma"] = {m: top * (l2m[m] if isinstance(l2m, dict) else l2m)
                                  for m in sim_params["meas_types"]}
    if sim_flags.get("irf_convolution", None) is not None:
        sim_flags["IRF_tables"] = load_irf_tables(sim_flags["irf_convolution"], irf_dir)
    else:
        sim_flags["IRF_tables"] = None
    simulate(e_data, P, X, param_info, dict(sim_params), init_params, sim_flags, logger=logger,
             evaluator=evaluator, comm=comm)
    return N, P, X


def export(out_filename, P, X, logger=None):
    """*_P.npy / *_X.npy next to out_filename (dense_sampling.py:314-329)."""
    head, base = os.path.dirname(out_filename), os.path.basename(out_filename)
    os.makedirs(head, exist_ok=True)
    np.save(os.path.join(head, f"{base}_P.npy"), P)
    np.save(os.path.join(head, f"{base}_X.npy"), X)
