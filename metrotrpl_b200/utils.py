"""Standalone helpers of the reference's utils.py used around the hot path (utils.py:5-40)."""
import numpy as np


def search_c_grps(c_grps, i):
    """First member of the constraint group that contains i, else i (utils.py:5-14)."""
    for grp in c_grps:
        if i in grp:
            return grp[0]
    return i


def set_min_y(sol, vals, scale_shift):
    """Host twin of the in-kernel floor (utils.py:16-32); kept for callers that hold curves."""
    min_y = 10 ** min(vals - scale_shift)
    first = np.searchsorted(-sol, -min_y)
    sol[first:] = min_y
    return sol, min_y, len(sol[first:])


def unpack_simpar(sim_info, i):
    return sim_info["lengths"][i], sim_info["nx"][i], sim_info["meas_types"][i]
