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
    with np.errstate(over="ignore"):        # both branches of the where() are evaluated
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



def _approve_rows(new_states, shared_fields):
    """Vectorised approve_move: True where a (log-scaled) proposal passes every check."""
    order = shared_fields["names"]
    prior = shared_fields["prior_dist"]
    idx = shared_fields["_param_indexes"]
    do_log = np.asarray(shared_fields["do_log"], dtype=bool)
    active = np.asarray(shared_fields["active"], dtype=bool)
    linear = np.where(do_log[None, :], 10 ** new_states, new_states)
    lo = np.array([prior[n][0] for n in order], dtype=float)
    hi = np.array([prior[n][1] for n in order], dtype=float)
    inside = (lo[None, :] < linear) & (linear < hi[None, :])
    ok = np.all(inside | ~active[None, :], axis=1)
    if "p0" in order and "n0" in order:
        ok &= new_states[:, idx["p0"]] > new_states[:, idx["n0"]]
    if "tauN" in order and "tauP" in order:
        ltn = new_states[:, idx["tauN"]]
        ltp = new_states[:, idx["tauP"]]
        if not do_log[idx["tauN"]]:
            ltn = np.log10(ltn)
        if not do_log[idx["tauP"]]:
            ltp = np.log10(ltp)
        ok &= np.abs(ltn - ltp) <= 2
    return ok


def _advance_keeping_buffer(bitgen, n):
    """PCG64.advance() clears the generator's buffered half-word (`has_uint32` / `uinteger`).  The
    reference never advances: a pending 32-bit half left by an `integers()` swap draw
    (metropolis.py:71) is consumed by the next one, so the buffer is carried across the jump."""
    before = bitgen.state
    bitgen.advance(n)
    after = bitgen.state
    after["has_uint32"] = before["has_uint32"]
    after["uinteger"] = before["uinteger"]
    bitgen.state = after


MAX_LOGGED_FAILS = 8       # include/metrotrpl_b200.h TRPL_MAX_LOGGED_FAILS
_native = {"lib": None, "tried": False}


def _native_lib():
    """The C entry point trpl_make_trial_moves of the CUDA library (host code), or None."""
    if not _native["tried"]:
        _native["tried"] = True
        try:
            from . import _capi
            _native["lib"] = _capi.load_library()
        except Exception:
            _native["lib"] = None
    return _native["lib"]


def _make_trial_moves_native(lib, cur, trial_moves, shared_fields, RNG, logger):
    """All chains in one C call (csrc/proposals.h): same generator stream, same arithmetic."""
    import ctypes as C
    n_chains, n_par = cur.shape
    order = shared_fields["names"]
    idx = shared_fields["_param_indexes"]
    prior = shared_fields["prior_dist"]
    do_log = np.ascontiguousarray(shared_fields["do_log"], dtype=np.uint8)
    active = np.ascontiguousarray(shared_fields["active"], dtype=np.uint8)
    lo = np.array([prior[n][0] for n in order], dtype=np.float64)
    hi = np.array([prior[n][1] for n in order], dtype=np.float64)
    st = RNG.bit_generator.state["state"]
    mask64 = (1 << 64) - 1
    pcg_state = (C.c_uint64 * 2)(st["state"] >> 64, st["state"] & mask64)
    pcg_inc = (C.c_uint64 * 2)(st["inc"] >> 64, st["inc"] & mask64)
    dl = np.asarray(shared_fields["do_log"], dtype=bool)
    cur = np.ascontiguousarray(np.where(dl[None, :], np.log10(cur), cur), dtype=np.float64)
    moves = np.ascontiguousarray(np.broadcast_to(trial_moves, cur.shape), dtype=np.float64)
    proposals = np.empty_like(cur)
    u = np.empty(n_chains)
    n_failed = np.zeros(n_chains, dtype=np.int32)
    masks = np.zeros((n_chains, MAX_LOGGED_FAILS), dtype=np.uint32)
    n_draws = C.c_int64(0)
    dp = C.POINTER(C.c_double)
    mu = shared_fields.get("do_mu_constraint", None)
    if mu is not None:
        # The reference draws the ambipolar mobility from the GLOBAL np.random stream, one draw per
        # attempt.  Enough uniforms for the worst case are pre-drawn, the C side reports how many it
        # used, and the global stream is then rewound and advanced by exactly that many.
        np_state = np.random.get_state()
        max_tries = MAX_PROPOSALS if shared_fields.get("hard_bounds", 0) else 1
        ambi_u = np.random.random_sample(n_chains * max_tries)
        ambi_lo, ambi_hi = float(mu[0] - mu[1]), float(mu[0] + mu[1])
        i_mun, i_mup = idx["mu_n"], idx["mu_p"]
    else:
        ambi_u = np.zeros(1)
        ambi_lo = ambi_hi = 0.0
        i_mun = i_mup = -1
    n_ambi_used = C.c_int32(0)
    mu_arg = np.zeros(n_chains)
    rc = lib.trpl_make_trial_moves(
        n_chains, n_par, cur.ctypes.data_as(dp), moves.ctypes.data_as(dp),
        do_log.ctypes.data_as(C.POINTER(C.c_uint8)), active.ctypes.data_as(C.POINTER(C.c_uint8)),
        lo.ctypes.data_as(dp), hi.ctypes.data_as(dp),
        idx["p0"] if "p0" in order and "n0" in order else -1, idx["n0"] if "p0" in order and "n0" in order else -1,
        idx["tauN"] if "tauN" in order and "tauP" in order else -1,
        idx["tauP"] if "tauN" in order and "tauP" in order else -1,
        1 if shared_fields.get("hard_bounds", 0) else 0, MAX_PROPOSALS, pcg_state, pcg_inc,
        proposals.ctypes.data_as(dp), u.ctypes.data_as(dp), C.byref(n_draws),
        n_failed.ctypes.data_as(C.POINTER(C.c_int32)), masks.ctypes.data_as(C.POINTER(C.c_uint32)),
        i_mun, i_mup, ambi_lo, ambi_hi, ambi_u.ctypes.data_as(dp), int(ambi_u.size), C.byref(n_ambi_used),
        mu_arg.ctypes.data_as(dp))
    if mu is not None:
        np.random.set_state(np_state)
        if n_ambi_used.value:
            np.random.random_sample(n_ambi_used.value)
        with np.errstate(invalid="ignore"):
            proposals[:, i_mup] = [np.log10(x) for x in mu_arg]     # NumPy's scalar log10, as the reference
    if rc != 0:
        raise RuntimeError(lib.trpl_last_error().decode())
    _advance_keeping_buffer(RNG.bit_generator, n_draws.value)
    proposals = np.where(dl[None, :], 10 ** proposals, proposals)
    if logger is not None and n_failed.any():
        # The reference warns once per failed attempt ("Failed checks: [...]"); with hundreds of hot
        # chains that is the most expensive thing the host does in an iteration.  One line per
        # iteration with the same information, counted per check.
        counts = {}
        logged = 0
        for m in np.nonzero(n_failed)[0]:
            for k in range(min(int(n_failed[m]), MAX_LOGGED_FAILS)):
                bits = int(masks[m, k])
                logged += 1
                for i in range(n_par):
                    if bits >> i & 1:
                        counts[f"{order[i]}_size"] = counts.get(f"{order[i]}_size", 0) + 1
                if bits >> 30 & 1:
                    counts["p0_greater"] = counts.get("p0_greater", 0) + 1
                if bits >> 31 & 1:
                    counts["tn_tp_close"] = counts.get("tn_tp_close", 0) + 1
        logger.warning(f"Failed checks: {int(n_failed.sum())} attempts of {int((n_failed > 0).sum())} chains "
                       f"rejected; per check (first {MAX_LOGGED_FAILS} attempts of a chain): {counts}")
    return proposals, u


def make_trial_moves(current_states, trial_moves, shared_fields, RNG, logger=None, native=True):
    """All chains' proposals and acceptance draws of one iteration.

    Consumes the generator exactly as the reference's serial loop does (metropolis.py:118-127):
    for each chain in turn its proposal draws (one RNG.random(n_params) per attempt, retried under
    hard bounds), then its acceptance draw.  The common case - first attempt admissible - is done for
    all remaining chains with one matrix draw; a chain that needs retries is replayed one draw at a
    time from the exact generator position (PCG64 `advance`), then the matrix draw resumes.

    Returns (proposals [n_chains, n_params], u [n_chains]).
    """
    # contiguous copy: NumPy's log10 / power take different code paths (SIMD or scalar, last-bit
    # different) for contiguous and strided operands, and states[:, :, k-1] is a strided view
    cur = np.ascontiguousarray(current_states, dtype=float)
    n_chains, n_par = cur.shape
    bitgen = RNG.bit_generator
    can_batch = shared_fields.get("do_mu_constraint", None) is None and hasattr(bitgen, "advance")
    if native and hasattr(bitgen, "advance") and n_par <= 30 and bitgen.state.get("bit_generator") == "PCG64":
        lib = _native_lib()
        if lib is not None and hasattr(lib, "trpl_make_trial_moves"):
            return _make_trial_moves_native(lib, cur, trial_moves, shared_fields, RNG, logger)
    proposals = np.empty_like(cur)
    u = np.empty(n_chains)
    do_log = np.asarray(shared_fields["do_log"], dtype=bool)
    hard = bool(shared_fields.get("hard_bounds", 0))
    m = 0
    while m < n_chains:
        if not can_batch:
            proposals[m] = make_trial_move(cur[m], trial_moves[m], shared_fields, RNG, logger)
            u[m] = RNG.random()
            m += 1
            continue
        state0 = bitgen.state
        rest = n_chains - m
        draws = RNG.random((rest, n_par + 1))
        logcur = np.where(do_log[None, :], np.log10(cur[m:]), cur[m:])
        new = logcur + trial_moves[m:] * (2 * draws[:, :n_par] - 1)
        ok = _approve_rows(new, shared_fields) if hard else np.ones(rest, dtype=bool)
        n_ok = rest if ok.all() else int(np.argmin(ok))
        proposals[m:m + n_ok] = np.where(do_log[None, :], 10 ** new[:n_ok], new[:n_ok])
        u[m:m + n_ok] = draws[:n_ok, n_par]
        m += n_ok
        if m < n_chains:
            # chain m needs retries: rewind to just before its first attempt and replay serially
            bitgen.state = state0
            _advance_keeping_buffer(bitgen, n_ok * (n_par + 1))
            proposals[m] = make_trial_move(cur[m], trial_moves[m], shared_fields, RNG, logger)
            u[m] = RNG.random()
            m += 1
    return proposals, u
