"""CPU tier: proposals, acceptance draws, swap decisions and whole chains against golden records of
the UNMODIFIED reference's metro(serial_fallback=True) (tools/make_golden.py gen_chains ->
tests/golden/chains.npz; metropolis.py:42-137, trial_move_generation.py:54-96 of the reference).
The likelihood backend is the host lock-step build of the kernel source (tests/emu)."""
import copy
import os
import tempfile

import numpy as np
import pytest

from metrotrpl_b200.metropolis import metro
from metrotrpl_b200.trial_move_generation import make_trial_moves
from tests.test_metropolis_batched import emu_factory, small_problem

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chains.npz"))
N_CHAINS, N_ITERS = 4, 20
# Visited states agree to rounding, not bit for bit: the reference's solve() scales the state it is
# given by the unit conversions IN PLACE and divides them back out afterwards
# (forward_solver.py:119-126, 200), and it is given a view of the History, so every evaluated state
# is perturbed by an ulp of that round trip (4.4e-29 becomes 4.399999999999999e-29 at iteration 0).
# The mirror leaves the state untouched.  Proposals from identical inputs ARE bit-identical
# (test_proposals_and_acceptance_draws_bit_exact), and every accept/swap decision is identical.
STATE_RTOL = 1e-13


def set_pcg(rng, words):
    w = [int(x) for x in words]
    st = rng.bit_generator.state
    st["state"]["state"] = (w[0] << 64) | w[1]
    st["state"]["inc"] = (w[2] << 64) | w[3]
    st["has_uint32"], st["uinteger"] = w[4], w[5]
    rng.bit_generator.state = st


def pcg_words(rng):
    st = rng.bit_generator.state
    m = (1 << 64) - 1
    return [st["state"]["state"] >> 64, st["state"]["state"] & m, st["state"]["inc"] >> 64,
            st["state"]["inc"] & m, st["has_uint32"], st["uinteger"]]


def variant(tag, tmp):
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp, n_chains=N_CHAINS, num_iters=N_ITERS)
    MCMC["checkpoint_freq"] = N_ITERS
    narrow = tag in ("bounds", "free", "notemper")
    MCMC["hard_bounds"] = 0 if tag == "free" else 1
    if narrow:
        param_info["prior_dist"]["p0"] = (2e15, 4e15)
        param_info["prior_dist"]["tauN"] = (400, 650)
        param_info["prior_dist"]["Sf"] = (5, 20)
    if tag in ("mu", "notemper"):
        MCMC["temper_freq"] = 1000
    if tag == "mu":
        MCMC["do_mu_constraint"] = (20.0, 3.0)
        param_info["active"]["mu_n"] = 1
        param_info["active"]["mu_p"] = 1
    e_data = (e_data[0], [v for v in G[f"{tag}_vals"]], e_data[2])     # the measurement the golden run saw
    return sim_info, ini, e_data, MCMC, param_info


def shared_fields_of(tag):
    from metrotrpl_b200.sim_utils import Ensemble
    with tempfile.TemporaryDirectory() as tmp:
        sim_info, ini, e_data, MCMC, param_info = variant(tag, tmp)
        ens = Ensemble(param_info, sim_info, MCMC, N_ITERS, False)
    return ens.ensemble_fields


def rolls_per_iteration(tag, k):
    swaps = tag in ("bounds", "free") and k % 2 == 0
    return N_CHAINS + (N_CHAINS - 1 if swaps else 0)


@pytest.mark.parametrize("tag", ["bounds", "free", "notemper", "mu"])
@pytest.mark.parametrize("native", [True, False])
def test_proposals_and_acceptance_draws_bit_exact(tag, native):
    """Every proposal of every chain and iteration, the acceptance draw that follows it and the
    generator state left behind, from the recorded generator state and current states: bit-exact,
    through the native (C) path and through the NumPy path."""
    sf = shared_fields_of(tag)
    rng = np.random.default_rng(0)
    roll = 0
    if tag == "mu":
        np.random.seed(1234)          # the reference's mu constraint draws from the global np.random
    for k in range(1, N_ITERS):
        lo = (k - 1) * N_CHAINS
        set_pcg(rng, G[f"{tag}_prop_rng"][lo])
        cur = G[f"{tag}_prop_in"][lo:lo + N_CHAINS]
        moves = G[f"{tag}_prop_move"][lo:lo + N_CHAINS]
        props, u = make_trial_moves(cur, moves, sf, rng, None, native=native)
        np.testing.assert_array_equal(props, G[f"{tag}_prop_out"][lo:lo + N_CHAINS])
        np.testing.assert_array_equal(u, G[f"{tag}_u"][roll:roll + N_CHAINS])
        n_rolls = rolls_per_iteration(tag, k)
        if n_rolls == N_CHAINS and k + 1 < N_ITERS:       # no swap round: the next proposal starts here
            assert pcg_words(rng) == [int(x) for x in G[f"{tag}_prop_rng"][lo + N_CHAINS]]
        roll += n_rolls
    assert roll == len(G[f"{tag}_u"])


def run_ours(tag, factory=emu_factory, **kw):
    """factory=None: the CUDA evaluator (GPU tier, tests/test_gpu_drivers.py)."""
    with tempfile.TemporaryDirectory() as tmp:
        sim_info, ini, e_data, MCMC, param_info = variant(tag, tmp)
        if tag == "mu":
            np.random.seed(1234)
        return metro(sim_info, ini, e_data, MCMC, param_info, export_path="out.pik", evaluator_factory=factory,
                     install_signal_handlers=False, **kw)


@pytest.mark.parametrize("tag", ["notemper", "mu"])
def test_whole_chains_without_swaps_equal_the_reference(tag):
    ms = run_ours(tag)
    np.testing.assert_array_equal(ms.H.accept, G[f"{tag}_accept"])            # every decision
    np.testing.assert_allclose(ms.H.states, G[f"{tag}_states"], rtol=STATE_RTOL, atol=0)
    np.testing.assert_allclose(ms.H.loglikelihood, G[f"{tag}_logll"], rtol=5e-3, atol=2e-3)   # the reference runs LSODA at its default tolerances (curves to ~1e-5)
    rng = np.random.default_rng(0)
    rng.bit_generator.state = ms.random_state
    assert pcg_words(rng) == [int(x) for x in G[f"{tag}_final_rng"]]


@pytest.mark.parametrize("tag", ["bounds", "free"])
def test_tempering_chains_equal_the_reference_serial_path(tag):
    """With the reference's swap aliasing reproduced on request, whole tempering chains - states,
    decisions, swap counts, generator - equal the reference's serial path; with the default
    (correct) swap they agree up to the first accepted swap, where the two chains involved then
    hold each other's states instead of both holding the upper one's."""
    ms = run_ours(tag, reference_swap_aliasing=True)
    np.testing.assert_array_equal(ms.H.accept, G[f"{tag}_accept"])
    np.testing.assert_allclose(ms.H.states, G[f"{tag}_states"], rtol=STATE_RTOL, atol=0)
    np.testing.assert_array_equal(ms.H.swap_accept, G[f"{tag}_swap_accept"])
    np.testing.assert_array_equal(ms.H.swap_attempts, G[f"{tag}_swap_attempts"])
    np.testing.assert_allclose(ms.H.loglikelihood, G[f"{tag}_logll"], rtol=5e-3, atol=2e-3)
    rng = np.random.default_rng(0)
    rng.bit_generator.state = ms.random_state
    assert pcg_words(rng) == [int(x) for x in G[f"{tag}_final_rng"]]
    # default behaviour: identical until the first accepted swap
    ok = G[f"{tag}_swap_ok"]
    first = int(np.argmax(ok))
    assert ok[first]
    k0, i0 = int(G[f"{tag}_swap_k"][first]), int(G[f"{tag}_swap_i"][first])
    ms2 = run_ours(tag)
    np.testing.assert_allclose(ms2.H.states[:, :, :k0], G[f"{tag}_states"][:, :, :k0], rtol=STATE_RTOL, atol=0)
    np.testing.assert_array_equal(ms2.H.accept[:, :k0 + 1], G[f"{tag}_accept"][:, :k0 + 1])
    if not np.any(ok[first + 1:][G[f"{tag}_swap_k"][first + 1:] == k0]):      # the only accepted swap of that round
        ref_k0 = G[f"{tag}_states"][:, :, k0]
        np.testing.assert_allclose(ms2.H.states[i0, :, k0], ref_k0[i0], rtol=STATE_RTOL)   # lower chain: the upper chain's state
        assert not np.allclose(ms2.H.states[i0 + 1, :, k0], ref_k0[i0 + 1], rtol=1e-6)     # upper chain: the lower one's, not its own


# ---- configs[0]: one chain on the reference's real measurement ---------------------------------
def real_chain_problem(tmp):
    from tests import parity_cases as pc
    import bench
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "staub_real3.npz"))
    n_t = g["n_t"]
    e_data = ([g["t"][m, :n_t[m]] for m in range(3)], [g["vals"][m, :n_t[m]] for m in range(3)],
              [g["uncs"][m, :n_t[m]] for m in range(3)])
    names = list(bench.NAMES)
    sim_info = {"lengths": [311.0] * 3, "nx": [128] * 3, "meas_types": ["TRPL"] * 3, "num_meas": 3}
    active = {n: int(n not in ("n0", "eps", "Tm", "m")) for n in names}
    param_info = {"names": names, "active": active, "unit_conversions": dict(zip(names, bench.UNITS)),
                  "do_log": {n: 1 for n in names},
                  "prior_dist": {n: ((lo, hi) if active[n] else (0, np.inf)) for n, lo, hi in zip(names, bench.LO, bench.HI)},
                  "init_guess": dict(zip(names, bench.GUESS)), "trial_move": {n: 0.1 for n in names}}
    param_info["prior_dist"]["m"] = (-np.inf, np.inf)
    MCMC = {"init_cond_path": "real_staub_input.csv", "measurement_path": "real_staub_aug_corr_renoised.csv",
            "output_path": tmp, "num_iters": 16, "solver": ("solveivp",), "model": "std", "ini_mode": "density",
            "log_y": 1, "checkpoint_freq": 16, "hard_bounds": 1, "rtol": None, "atol": None,
            "model_uncertainty": {"TRPL": 1.0}}
    return sim_info, g["ini"].copy(), e_data, MCMC, param_info


def check_real_chain(ms):
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_real.npz"))
    np.testing.assert_array_equal(ms.H.accept, G["accept"])                       # every decision
    np.testing.assert_allclose(ms.H.states, G["states"], rtol=STATE_RTOL, atol=0)
    # the reference runs LSODA at its default tolerances: its own likelihoods are good to ~1e-4 here
    np.testing.assert_allclose(ms.H.loglikelihood, G["logll"], rtol=2e-3, atol=2e-3)


def test_single_chain_on_the_real_staub_data_equals_the_reference():
    """configs[0]: Inputs/mcmc0.txt's parameter set on the reference's real measurement (three curves,
    nx = 128), one chain, against the chain the unmodified reference's metro(serial_fallback=True)
    walks (tools/make_golden.py gen_chain_real)."""
    with tempfile.TemporaryDirectory() as tmp:
        sim_info, ini, e_data, MCMC, param_info = real_chain_problem(tmp)
        ms = metro(sim_info, ini, e_data, MCMC, param_info, export_path="out.pik", evaluator_factory=emu_factory,
                   install_signal_handlers=False)
    check_real_chain(ms)
