"""Proposal generation (host side), mirroring trial_move_generation.py of the reference.

The proposals themselves are cheap; what matters for the GPU is that all chains' proposals of one
iteration are generated first and evaluated together (metropolis.py here).
"""
import numpy as np

from .sim_utils import MAX_PROPOSALS


def approve_move(new_state, shared_fields):
    """Names of the checks a proposed (log-scaled) state fails (trial_move_generation.py:4-52)."""
    order = shared_fields["names"]
    prior = shared_fields["prior_dist"]
    idx = shared_fields["_param_indexes"]
    do_log = shared_fields["do_log"]
    failed = []
    linear = np.where(do_log, 10 ** new_state, new_state)
    for i, name in enumerate(order):
        if not shared_fields["active"][i]:
            continue
        lo, hi = prior[name][0], prior[name][1]
        if not (lo < linear[i] < hi):
            failed.append(f"{name}_size")
    if "p0" in order and "n0" in order:                      # p-type by definition
        if not (new_state[idx["p0"]] > new_state[idx["n0"]]):
            failed.append("p0_greater")
    if "tauN" in order and "tauP" in order:                  # within two orders of magnitude
        ltn = new_state[idx["tauN"]]
        ltp = new_state[idx["tauP"]]
        if not do_log[idx["tauN"]]:
            ltn = np.log10(ltn)
        if not do_log[idx["tauP"]]:
            ltp = np.log10(ltp)
        if not (np.abs(ltn - ltp) <= 2):
            failed.append("tn_tp_close")
    return failed


def make_trial_move(current_state, trial_move, shared_fields, RNG, logger=None):
    """Uniform box displacement around the current state (trial_move_generation.py:54-96).

    Consumes the generator exactly as the reference does: one RNG.random(n_params) per attempt,
    up to MAX_PROPOSALS attempts when hard_bounds is set.
    """
    cur = np.array(current_state, dtype=float)
    do_log = shared_fields["do_log"]
    cur = np.where(do_log, np.log10(cur), cur)
    mu_constraint = shared_fields.get("do_mu_constraint", None)
    max_tries = MAX_PROPOSALS if shared_fields.get("hard_bounds", 0) else 1
    new_state = np.array(cur)
    for _ in range(max_tries):
        new_state = cur + trial_move * (2 * RNG.random(cur.shape) - 1)
        if mu_constraint is not None:
            ambi, ambi_std = mu_constraint[0], mu_constraint[1]
            new_ambi = np.random.uniform(ambi - ambi_std, ambi + ambi_std)
            i_n = shared_fields["_param_indexes"]["mu_n"]
            i_p = shared_fields["_param_indexes"]["mu_p"]
            new_state[i_p] = np.log10((2 / new_ambi - 1 / 10 ** new_state[i_n]) ** -1)
        failed = approve_move(new_state, shared_fields)
        if not failed:
            break
        if logger is not None:
            logger.warning(f"Failed checks: {failed}")
    return np.where(do_log, 10 ** new_state, new_state)
