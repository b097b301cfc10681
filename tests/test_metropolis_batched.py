"""CPU tier: host logic of the batched Metropolis / dense-sampling drivers, including the N > 1
path under torch.distributed (gloo, world_size 2).  The likelihood backend is injected: the host
lock-step build of the kernel source (tests/emu) stands in for the GPU here."""
import copy
import os
import pickle
import sys
import tempfile

import numpy as np
import pytest

from metrotrpl_b200 import _capi
from metrotrpl_b200.metropolis import metro, roll_acceptance
from metrotrpl_b200.trial_move_generation import approve_move, make_trial_move
from metrotrpl_b200.sim_utils import Ensemble, History
from oracle import trpl_oracle as orc

NAMES = ["n0", "p0", "mu_n", "mu_p", "ks", "Cn", "Cp", "Sf", "Sb", "tauN", "tauP", "eps", "Tm", "m"]
UNITS = {"n0": 1e-21, "p0": 1e-21, "mu_n": 1e5, "mu_p": 1e5, "ks": 1e12, "Cn": 1e33, "Cp": 1e33,
         "Sf": 1e-2, "Sb": 1e-2}
GUESS = {"n0": 1e8, "p0": 3e15, "mu_n": 20, "mu_p": 20, "ks": 4.8e-11, "Cn": 4.4e-29, "Cp": 4.4e-29,
         "Sf": 10, "Sb": 10, "tauN": 511, "tauP": 871, "eps": 10, "Tm": 300, "m": 1}


def small_problem(tmp, n_chains=4, num_iters=6):
    nx = 32
    t = np.linspace(0, 40, 9)
    x = (np.arange(nx) + 0.5) * (311 / nx)
    ini = np.array([2e16 * np.exp(-x / 100.0), 2e17 * np.exp(-x / 100.0)])
    sim_info = {"lengths": [311, 311], "nx": [nx, nx], "meas_types": ["TRPL", "TRPL"], "num_meas": 2}
    idx = {n: i for i, n in enumerate(NAMES)}
    units = np.array([UNITS.get(n, 1) for n in NAMES], dtype=float)
    state = np.array([GUESS[n] for n in NAMES], dtype=float)
    vals = [np.log10(orc.simulate(ini[m], orc.Grid(311, nx, t, 4), state, idx, units=units)) + 0.05
            for m in range(2)]
    uncs = [np.full(len(t), 0.02)] * 2
    param_info = {"names": list(NAMES), "active": {n: int(n in ("p0", "ks", "tauN", "tauP", "Sf")) for n in NAMES},
                  "unit_conversions": dict(UNITS), "do_log": {n: 1 for n in NAMES},
                  "prior_dist": {n: (0, np.inf) for n in NAMES}, "init_guess": dict(GUESS),
                  "trial_move": {n: 0.05 for n in NAMES}}
    param_info["prior_dist"]["m"] = (-np.inf, np.inf)
    MCMC = {"init_cond_path": "x", "measurement_path": "y", "output_path": tmp, "num_iters": num_iters,
            "solver": ("solveivp",), "model": "std", "ini_mode": "density", "log_y": 1,
            "checkpoint_freq": 4, "hard_bounds": 1, "rtol": 1e-6, "atol": None,
            "model_uncertainty": {"TRPL": 0.05},
            "parallel_tempering": list(2.0 ** np.arange(n_chains)), "temper_freq": 2}
    return sim_info, ini, ([t, t], vals, uncs), MCMC, param_info


def emu_factory(shared_fields):
    from tests.emu import emu
    from metrotrpl_b200.trial_move_evaluation import PathCache

    class Ev:
        def __init__(self):
            sf = shared_fields
            self.prob = _capi.pack_problem(sf["_sim_info"], sf["_init_params"], sf["_times"], sf["_vals"],
                                           sf["_uncs"], model=sf["model"], ini_mode=sf["ini_mode"])
            self.ladder = np.asarray(sf["_T"], dtype=float)
            self.sf = sf

        def __call__(self, states, sigmas):
            sf = self.sf
            params = _capi.pack_params(states, sf["_param_indexes"], sf["units"], model=sf["model"])
            n = len(states)
            sig = np.array([[s[t] for t in sf["_sim_info"]["meas_types"]] for s in sigmas])
            aux = np.zeros((n, self.prob.n_meas, _capi.NAUX))
            aux[..., _capi.A_FLUENCE_MULT] = 1
            aux[..., _capi.A_ABSORB_MULT] = 1
            for k in range(3):
                aux[..., _capi.A_S2T0 + k] = sig ** 2
            opts = _capi.make_opts(sf["rtol"], sf["atol"], flags=_capi.OPT_LADDER)
            _, _, _, _, lad = emu.loglik_batch(self.prob, params, aux, opts, True, ladder=self.ladder)
            return lad.sum(axis=1)
    return Ev()


def run_metro(tmp, **kw):
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp)
    return metro(sim_info, ini, e_data, MCMC, param_info, export_path="out.pik",
                 evaluator_factory=emu_factory, install_signal_handlers=False, **kw)


def test_roll_acceptance_matches_reference_numbers():
    rng = np.random.default_rng(1)                      # Tests/test_metropolis.py:367-380
    assert all(roll_acceptance(rng, np.ones(100)))
    assert roll_acceptance(rng, np.ones(10000) * -1).sum() == 3635


def test_make_trial_move_respects_bounds_and_inactive():
    sim_info, ini, e_data, MCMC, param_info = small_problem(tempfile.mkdtemp())
    param_info["prior_dist"]["tauN"] = (500, 520)
    ens = Ensemble(param_info, sim_info, MCMC, 5)
    sf = ens.ensemble_fields
    rng = np.random.default_rng(3)
    cur = ens.H.states[0, :, 0]
    for _ in range(20):
        new = make_trial_move(cur, sf["base_trial_move"], sf, rng)
        assert 500 < new[sf["_param_indexes"]["tauN"]] < 520
        inactive = ~sf["active"]
        np.testing.assert_allclose(new[inactive], cur[inactive], rtol=1e-12)
    bad = np.log10(cur.copy())
    bad[sf["_param_indexes"]["tauN"]] = np.log10(600)
    assert "tauN_size" in approve_move(bad, sf)


def test_batched_proposals_consume_the_generator_like_the_serial_loop():
    from metrotrpl_b200.trial_move_generation import make_trial_moves
    sim_info, ini, e_data, MCMC, param_info = small_problem(tempfile.mkdtemp(), n_chains=12)
    param_info["prior_dist"]["tauN"] = (400, 650)          # tight box: some chains need retries
    param_info["prior_dist"]["p0"] = (1e15, 1e16)
    ens = Ensemble(param_info, sim_info, MCMC, 5)
    sf = ens.ensemble_fields
    T = np.asarray(sf["_T"], dtype=float)
    cur = np.repeat(ens.H.states[:, :, 0][:1], 12, axis=0) * (1 + 0.01 * np.arange(12))[:, None]
    moves = np.sqrt(T)[:, None] * sf["base_trial_move"][None, :]
    for seed in range(5):
        r1 = np.random.default_rng(seed)
        r2 = np.random.default_rng(seed)
        ref_p, ref_u = [], []
        for m in range(12):
            ref_p.append(make_trial_move(cur[m], moves[m], sf, r1))
            ref_u.append(r1.random())
        got_p, got_u = make_trial_moves(cur, moves, sf, r2)
        np.testing.assert_array_equal(got_p, np.array(ref_p))
        np.testing.assert_array_equal(got_u, np.array(ref_u))
        assert r1.random() == r2.random()                  # same generator position afterwards


def test_metro_runs_checkpoints_and_is_deterministic():
    tmp = tempfile.mkdtemp()
    a = run_metro(tmp)
    assert a.H.states.shape == (4, len(NAMES), 6)
    assert np.all(np.isfinite(a.H.loglikelihood))
    assert a.H.accept[:, 1:].sum() > 0
    assert a.H.swap_attempts.sum() == 3 * 2          # temper_freq 2 -> k = 2, 4; n_chains-1 attempts each
    assert os.path.exists(os.path.join(tmp, "out.pik"))
    with open(os.path.join(tmp, "out.pik"), "rb") as f:
        b = pickle.load(f)
    np.testing.assert_array_equal(a.H.states, b.H.states)
    c = run_metro(tempfile.mkdtemp())
    np.testing.assert_array_equal(a.H.states, c.H.states)
    np.testing.assert_array_equal(a.H.loglikelihood, c.H.loglikelihood)
    # the first column is the likelihood of the initial guess at each chain's own temperature:
    # check it against the oracle (SciPy LSODA) directly
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp)
    idx = {n: i for i, n in enumerate(NAMES)}
    units = np.array([UNITS.get(n, 1) for n in NAMES], dtype=float)
    state = np.array([GUESS[n] for n in NAMES], dtype=float)
    for m, T in enumerate(MCMC["parallel_tempering"]):
        ll, _ = orc.state_loglik(state, sim_info, ini, e_data[0], e_data[1], e_data[2], idx, units,
                                 {"TRPL": 0.05}, T=T, rtol=1e-10, atol=1e-16)
        assert abs(a.H.loglikelihood[m, 0] / ll - 1) < 1e-4


def _worker(rank, world, port, tmp, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from metrotrpl_b200.parallel import Comm
    comm = Comm(backend="gloo")
    ms = run_metro(os.path.join(tmp, f"r{rank}"), comm=comm)
    q.put((rank, ms.H.states, ms.H.loglikelihood, ms.H.swap_accept))
    comm.barrier()


def test_two_ranks_equal_one_rank_gloo():
    import torch.multiprocessing as mp
    tmp = tempfile.mkdtemp()
    one = run_metro(os.path.join(tmp, "single"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, states, ll, sw in got:
        np.testing.assert_array_equal(states, one.H.states)
        np.testing.assert_array_equal(ll, one.H.loglikelihood)
        np.testing.assert_array_equal(sw, one.H.swap_accept)


def test_dense_sampling_blocks_and_sharding():
    from metrotrpl_b200 import dense_sampling as ds
    from metrotrpl_b200.parallel import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    tmp = tempfile.mkdtemp()
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp)
    param_info["prior_dist"]["tauN"] = (100, 1000)
    param_info["prior_dist"]["p0"] = (1e15, 1e16)
    for n in NAMES:
        if n not in ("tauN", "p0"):
            param_info["active"][n] = 0
    flags = {"num_iters": 7, "log_y": 1, "likel2move_ratio": {"TRPL": 1.0}, "model": "std",
             "ini_mode": "density", "rtol": 1e-6}
    np.random.seed(0)
    calls = []

    def fake_eval(states):
        calls.append(len(states))
        return -np.log10(states[:, 9])      # any deterministic function of the sample

    N, P, X = ds.bayes(np.array([0]), None, ini, sim_info, e_data, flags, param_info, evaluator=fake_eval)
    assert X.shape == (7, len(NAMES)) and P.shape == (7,)
    assert np.all((X[:, 9] > 100) & (X[:, 9] < 1000))
    assert np.all(X[:, 0] == GUESS["n0"])
    np.testing.assert_allclose(P, -np.log10(X[:, 9]))
    assert sum(calls) == 7
    ds.export(os.path.join(tmp, "dense", "CPU0"), P, X)
    assert os.path.exists(os.path.join(tmp, "dense", "CPU0_P.npy"))


def test_history_extend_truncate():
    h = History(2, 5, ["a", "b"])
    h.extend(8)
    assert h.states.shape == (2, 2, 8) and h.accept.shape == (2, 8)
    h.extend(3)
    assert h.loglikelihood.shape == (2, 3)


def test_native_proposals_equal_the_python_mirror_bit_for_bit():
    """trpl_make_trial_moves (C, one call for all chains) against make_trial_moves' NumPy path:
    same proposals, same acceptance draws, same generator position, same warnings - with hot
    chains whose large moves need many retries under hard bounds."""
    import logging
    from metrotrpl_b200.trial_move_generation import make_trial_moves
    tmp = tempfile.mkdtemp()
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp, n_chains=64)
    param_info["prior_dist"].update({"p0": (1e14, 1e16), "ks": (1e-11, 1e-9), "tauN": (1, 1500),
                                     "tauP": (1, 3000), "Sf": (1e-4, 1e4)})
    ens = Ensemble(param_info, sim_info, MCMC, 4)
    sf = ens.ensemble_fields
    T = np.asarray(sf["_T"], dtype=float)
    moves = np.sqrt(np.minimum(T, 4000.0))[:, None] * sf["base_trial_move"][None, :] * 4
    cur = np.repeat(ens.H.states[:, :, 0][:1], len(T), axis=0)

    class Rec(logging.Handler):
        def __init__(self):
            super().__init__()
            self.msgs = []

        def emit(self, record):
            self.msgs.append(record.getMessage())

    outs = []
    for native in (True, False):
        rng = np.random.default_rng(77)
        log = logging.getLogger(f"moves{native}")
        log.setLevel(logging.WARNING)
        rec = Rec()
        log.addHandler(rec)
        steps = []
        state = cur.copy()
        for _ in range(5):
            p, u = make_trial_moves(state, moves, sf, rng, log, native=native)
            steps.append((p.copy(), u.copy()))
            state = p
        outs.append((steps, rng.bit_generator.state["state"]["state"], rec.msgs))
    (sa, ga, ma), (sb, gb, mb) = outs
    assert ga == gb
    for (pa, ua), (pb, ub) in zip(sa, sb):
        np.testing.assert_array_equal(pa, pb)
        np.testing.assert_array_equal(ua, ub)
    assert len(mb) > 20                      # the case does exercise the retry path
    # the native path writes one summary line per call; its attempt count is the number of lines the
    # serial path writes
    assert len(ma) == 5
    assert sum(int(m.split()[2]) for m in ma) == len(mb)


def _approve_fields(do_log, active):
    names = ["tauP", "tauN", "somethingelse"]
    prior = {"tauP": (0.1, np.inf), "tauN": (0.1, np.inf), "somethingelse": (-np.inf, np.inf)}
    return {"do_log": np.array(do_log, dtype=bool), "active": np.array(active, dtype=bool),
            "prior_dist": prior, "_param_indexes": {n: i for i, n in enumerate(names)}, "names": names,
            "hard_bounds": 1}


def _native_failed_names(state_in_sampler_scale, sf):
    """The checks the C implementation reports for one fixed state (move size zero: every attempt
    proposes the state itself)."""
    from metrotrpl_b200 import trial_move_generation as tmg
    import logging
    msgs = []

    class H(logging.Handler):
        def emit(self, record):
            msgs.append(record.getMessage())
    log = logging.getLogger("approve_native")
    log.setLevel(logging.WARNING)
    log.handlers = [H()]
    dl = sf["do_log"]
    lin = np.where(dl, 10 ** np.asarray(state_in_sampler_scale, dtype=float), state_in_sampler_scale)
    rng = np.random.default_rng(0)
    tmg.make_trial_moves(lin[None, :], np.zeros((1, 3)), sf, rng, log)
    if not msgs:
        return []
    import ast
    counts = ast.literal_eval(msgs[0].split("per check (first 8 attempts of a chain): ")[1])
    return sorted(counts)


def test_approve_move_reference_cases_python_and_native():
    """Tests/test_approve_move.py:15-80 (tauN/tauP within two decades, size limits, inactive
    parameters, parameters that are not log-scaled) - on the Python mirror and on the checks inside
    trpl_make_trial_moves."""
    for do_log, scale in (([1, 1, 1], np.log10), ([0, 0, 1], lambda x: np.array(x, dtype=float))):
        sf = _approve_fields(do_log, [1, 1, 1])

        def st(v):
            v = np.array(v, dtype=float)
            out = scale(v)
            if do_log == [0, 0, 1]:
                out[2] = np.log10(v[2])
            return out
        cases = [([511, 511e2, 1], []), ([511, 511e2 + 1, 1], ["tn_tp_close"]), ([0.11, 0.11, 1], []),
                 ([0.1, 0.11, 1], ["tauP_size"]), ([0.11, 0.1, 1], ["tauN_size"])]
        for vals, want in cases:
            got = approve_move(st(vals), sf)
            assert sorted(got) == want, (vals, got)
            if do_log == [1, 1, 1]:
                # the native path receives log-scaled states: same arithmetic, same verdicts
                assert _native_failed_names(st(vals), sf) == want, vals
    sf = _approve_fields([1, 1, 1], [0, 0, 1])
    assert approve_move(np.log10([0.11, 0.1, 1]), sf) == []
    assert _native_failed_names(np.log10([0.11, 0.1, 1]), sf) == []


def _dense_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from metrotrpl_b200 import dense_sampling as ds
    from metrotrpl_b200.parallel import Comm
    comm = Comm(backend="gloo")
    sim_info, ini, e_data, MCMC, param_info = small_problem(tempfile.mkdtemp())
    param_info["prior_dist"]["tauN"] = (100, 1000)
    param_info["prior_dist"]["p0"] = (1e15, 1e16)
    for n in NAMES:
        if n not in ("tauN", "p0"):
            param_info["active"][n] = 0
    flags = {"num_iters": 9, "log_y": 1, "likel2move_ratio": {"TRPL": 1.0}, "model": "std",
             "ini_mode": "density", "rtol": 1e-6}
    np.random.seed(100 + rank)           # every rank's global generator is in a DIFFERENT state
    N, P, X = ds.bayes(np.array([0]), None, ini, sim_info, e_data, flags, param_info,
                       evaluator=lambda s: -np.log10(s[:, 9]) + np.log10(s[:, 1]), comm=comm)
    q.put((rank, P, X))
    comm.barrier()


def test_dense_grid_is_rank_zeros_on_every_rank_gloo():
    """bayes() draws its grid from the unseeded global np.random (as the reference does, which never
    shards): rank 0's grid is broadcast, so the gathered likelihoods belong to the X every rank returns."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_dense_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=600) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    (_, P0, X0), (_, P1, X1) = got
    np.testing.assert_array_equal(X0, X1)
    np.testing.assert_array_equal(P0, P1)
    np.testing.assert_allclose(P0, -np.log10(X0[:, 9]) + np.log10(X0[:, 1]))
