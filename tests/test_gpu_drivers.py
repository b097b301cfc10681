"""GPU tier: the batched callers of the path (eval_trial_move mirror, solve mirror, metro(),
dense sampling) through the CUDA library."""
import os
import tempfile

import numpy as np
import pytest

from metrotrpl_b200 import _capi
from metrotrpl_b200 import dense_sampling as ds
from metrotrpl_b200.forward_solver import solve
from metrotrpl_b200.metropolis import metro
from metrotrpl_b200.sim_utils import Grid
from metrotrpl_b200.trial_move_evaluation import eval_trial_move
from oracle import trpl_oracle as orc
from tests.test_metropolis_batched import GUESS, NAMES, UNITS, emu_factory, small_problem

pytestmark = pytest.mark.gpu


def test_solve_mirror_matches_reference_unit_test():
    """Tests/test_metropolis.py:93-190 (test_solve): high-injection radiative decay, PL and TRTS,
    undefined measurement / solver raise NotImplementedError, state is left untouched."""
    g = Grid(thickness=1000, nx=100, tSteps=np.linspace(0, 100, 1001), hmax=4)
    names = ["n0", "p0", "mu_n", "mu_p", "ks", "tauN", "tauP", "Cn", "Cp", "Sf", "Sb", "eps", "Tm"]
    uc = {"n0": 1e-21, "p0": 1e-21, "mu_n": 1e5, "mu_p": 1e5, "ks": 1e12, "Sf": 1e-2, "Sb": 1e-2}
    vals = {"n0": 0, "p0": 0, "mu_n": 0, "mu_p": 0, "ks": 1e-11, "Cn": 0, "Cp": 0, "tauN": 1e99,
            "tauP": 1e99, "Sf": 0, "Sb": 0, "Tm": 300, "eps": 1}
    idx = {n: i for i, n in enumerate(names)}
    state = [vals[n] for n in names]
    units = np.array([uc.get(n, 1) for n in names], dtype=float)
    init_dN = 1e20 * np.ones(g.nx)
    pl = solve(init_dN, g, state, idx, meas="TRPL", units=units, RTOL=1e-10, ATOL=1e-14)
    out_dN = 0.0009900990095719482
    expect = 1e-11 * 1e12 * out_dN ** 2 * 1000 * 1e23
    assert abs(pl[-1] / expect - 1) < 1e-7
    vals2 = dict(vals, mu_n=10, mu_p=10)
    state2 = [vals2[n] for n in names]
    trts = solve(init_dN, g, state2, idx, meas="TRTS", units=units)
    expect = orc.Q_C * (2 * 10 * 1e5) * 0.0009900986886696803 * 1000 * 1e9
    assert abs(trts[-1] / expect - 1) < 1e-6
    assert state2 == [vals2[n] for n in names]
    with pytest.raises(NotImplementedError):
        solve(init_dN, g, state, idx, meas="something else")
    with pytest.raises(NotImplementedError):
        solve(init_dN, g, state, idx, meas="TRPL", solver=("somethign else",))
    # Tests/test_metropolis.py:192-251 (test_solve_depletion): a raised Grid.min_y truncates the tail
    vals3 = dict(vals, mu_n=1, mu_p=1, ks=2e-10)
    st3 = [vals3[n] for n in names]
    pl0 = 2e-10 * (1e18) ** 2 * 1000e-7
    g.min_y = pl0 * 1e-2
    pl = solve(1e18 * np.ones(g.nx), g, st3, idx, meas="TRPL", units=units, RTOL=1e-10, ATOL=1e-14)
    assert min(pl) >= g.min_y
    np.testing.assert_equal(pl[-10:], g.min_y)
    # Tests/test_metropolis.py:314-365 (test_solve_iniPar): fluence mode == explicit profile
    g2 = Grid(thickness=1000, nx=100, tSteps=np.linspace(0, 100, 1001), hmax=4)
    prof = 1e15 * 6e4 * np.exp(-6e4 * g2.xSteps * 1e-7)
    a = solve(prof, g2, state, idx, units=units, RTOL=1e-10, ATOL=1e-14)
    b = solve([1e15, 6e4], g2, state, idx, units=units, ini_mode="fluence", RTOL=1e-10, ATOL=1e-14)
    np.testing.assert_allclose(a / a.max(), b / a.max(), atol=1e-7)


def test_eval_trial_move_mirror_and_ll_funcs():
    tmp = tempfile.mkdtemp()
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp)
    idx = {n: i for i, n in enumerate(NAMES)}
    units = np.array([UNITS.get(n, 1) for n in NAMES], dtype=float)
    state = np.array([GUESS[n] for n in NAMES], dtype=float)
    sf = {"_sim_info": sim_info, "_init_params": ini, "_times": e_data[0], "_vals": e_data[1],
          "_uncs": e_data[2], "_param_indexes": idx, "units": units, "model": "std",
          "ini_mode": "density", "hmax": 4, "rtol": 1e-8, "atol": None}
    uf = {"model_uncertainty": {"TRPL": 0.05}, "_T": 2.0, "_T_alt": (1.0, 4.0)}
    ll, funcs = eval_trial_move(state, uf, sf, None)
    for T in (2.0, 1.0, 4.0):
        want, per = orc.state_loglik(state, sim_info, ini, e_data[0], e_data[1], e_data[2], idx, units,
                                     {"TRPL": 0.05}, T=T, rtol=1e-10, atol=1e-16)
        got = sum(f(T) for f in funcs)
        assert abs(got / want - 1) < 1e-5
    assert abs(ll - sum(f(2.0) for f in funcs)) < 1e-12 * abs(ll)
    with pytest.raises(KeyError):
        funcs[0](3.0)


def test_metro_on_gpu_matches_host_lockstep_run():
    a = metro(*_args(tempfile.mkdtemp()), export_path="g.pik", install_signal_handlers=False)
    b = metro(*_args(tempfile.mkdtemp()), export_path="e.pik", install_signal_handlers=False,
              evaluator_factory=emu_factory)
    np.testing.assert_allclose(a.H.loglikelihood[:, 0], b.H.loglikelihood[:, 0], rtol=1e-6)
    # identical generator stream + likelihoods equal to rounding -> same decisions
    np.testing.assert_array_equal(a.H.accept, b.H.accept)
    np.testing.assert_allclose(a.H.states, b.H.states, rtol=1e-12)
    np.testing.assert_array_equal(a.H.swap_accept, b.H.swap_accept)


def _args(tmp):
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp)
    return sim_info, ini, e_data, MCMC, param_info


def test_dense_sampling_on_gpu_against_oracle():
    tmp = tempfile.mkdtemp()
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp)
    param_info["prior_dist"]["tauN"] = (100, 1000)
    param_info["prior_dist"]["p0"] = (1e15, 1e16)
    for n in NAMES:
        if n not in ("tauN", "p0"):
            param_info["active"][n] = 0
    flags = {"num_iters": 6, "log_y": 1, "likel2move_ratio": {"TRPL": 1.0}, "model": "std",
             "ini_mode": "density", "rtol": 1e-8}
    np.random.seed(1)
    N, P, X = ds.bayes(np.array([0]), None, ini, sim_info, e_data, flags, param_info)
    idx = {n: i for i, n in enumerate(NAMES)}
    units = np.array([UNITS.get(n, 1) for n in NAMES], dtype=float)
    sigma = 0.05 * 1.0
    for i in (0, 5):
        want, _ = orc.state_loglik(X[i], sim_info, ini, e_data[0], e_data[1], e_data[2], idx, units,
                                   {"TRPL": sigma}, rtol=1e-10, atol=1e-16)
        assert abs(P[i] / want - 1) < 1e-5


def test_dense_sampling_pipeline_equals_block_by_block():
    """simulate() alternates its blocks between two contexts (launches overlap); the result must be
    what one context gives block by block, bit for bit, whatever the block size."""
    tmp = tempfile.mkdtemp()
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp)
    names = param_info["names"]
    rng = np.random.default_rng(5)
    base = np.array([GUESS[n] for n in names], dtype=float)
    X = np.repeat(base[None, :], 37, axis=0)
    X[:, names.index("tauN")] = 10 ** rng.uniform(2, 3, 37)
    X[:, names.index("p0")] = 10 ** rng.uniform(15, 16, 37)
    flags = {"log_y": 1, "model": "std", "ini_mode": "density", "rtol": 1e-7, "current_sigma": {"TRPL": 0.05}}
    outs = []
    for block in (5, 16, 64):
        P = np.zeros(37)
        ds.simulate(e_data, P, X, param_info, dict(sim_info), ini, flags, block=block)
        outs.append(P)
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])
    from metrotrpl_b200.trial_move_evaluation import eval_trial_moves
    idx = {n: i for i, n in enumerate(names)}
    units = np.array([UNITS.get(n, 1) for n in names], dtype=float)
    sf = {"_sim_info": sim_info, "_init_params": ini, "_times": e_data[0], "_vals": e_data[1],
          "_uncs": e_data[2], "_param_indexes": idx, "units": units, "model": "std",
          "ini_mode": "density", "rtol": 1e-7, "atol": None}
    ref = eval_trial_moves(X, np.ones(37), {"TRPL": 0.05}, sf).logll
    np.testing.assert_array_equal(outs[0], ref)


def test_scale_fluence_and_absorption_factors_against_the_oracle():
    """`_s#`, `_f#`, `_a#` (trial_move_evaluation.py:38-60) end to end through PathCache and the
    kernel, fluence mode, with constraint groups; the oracle gets the factors applied to a fresh copy
    of the initial condition (the reference multiplies _init_params in place, so its factor
    compounds from call to call - documented deviation)."""
    from metrotrpl_b200.trial_move_evaluation import eval_trial_moves
    names = list(NAMES) + ["_s0", "_s1", "_f0", "_f2", "_a1"]
    units = np.array([UNITS.get(n, 1) for n in names], dtype=float)
    idx = {n: i for i, n in enumerate(names)}
    t = np.linspace(0, 60, 61)
    sim = {"lengths": [311.0, 500.0, 311.0], "nx": [64, 64, 64], "meas_types": ["TRPL"] * 3, "num_meas": 3}
    inis = [np.array([2e12, 6e4, 1.0]), np.array([6e12, 5e4, -1.0]), np.array([2e13, 6e4, 1.0])]
    rng = np.random.default_rng(3)
    vals = [17.0 - 0.02 * t + 0.01 * rng.standard_normal(61) for _ in range(3)]
    uncs = [np.full(61, 0.03)] * 3
    sf = {"_sim_info": sim, "_init_params": [i.copy() for i in inis], "_times": [t] * 3, "_vals": vals,
          "_uncs": uncs, "_param_indexes": idx, "units": units, "model": "std", "ini_mode": "fluence",
          "rtol": 1e-8, "atol": None, "hmax": 4,
          "scale_factor": (0.02, [0, 1, 2], [(0, 2)]),            # meas 0, 2 share _s0; meas 1 has _s1
          "fittable_fluences": (0.02, [0, 2], None),              # _f0, _f2
          "fittable_absps": (0.02, [1], None)}                    # _a1
    base = np.array([GUESS[n] for n in NAMES] + [3.0, 0.4, 1.7, 0.6, 1.3], dtype=float)
    other = base.copy()
    other[idx["_s0"]], other[idx["_f2"]], other[idx["_a1"]] = 0.7, 2.5, 0.8
    states = np.stack([base, other])
    res = eval_trial_moves(states, np.ones(2), {"TRPL": 1.0}, sf, want_curves=True)
    res2 = eval_trial_moves(states, np.ones(2), {"TRPL": 1.0}, sf, want_curves=True)
    np.testing.assert_array_equal(res.per_meas, res2.per_meas)    # no compounding across calls
    for k, st in enumerate(states):
        for m in range(3):
            ini = inis[m].copy()
            if m in (0, 2):
                ini[0] *= st[idx[f"_f{m}"]]
            if m == 1:
                ini[1] *= st[idx["_a1"]]
            shift = np.log10(st[idx["_s0"]] if m in (0, 2) else st[idx["_s1"]])
            g = orc.Grid(sim["lengths"][m], 64, t, 4)
            ref = orc.simulate(ini, g, st, idx, units=units, ini_mode="fluence", RTOL=1e-10, ATOL=1e-16)
            np.testing.assert_allclose(res.curves[k, 61 * m:61 * (m + 1)], ref, rtol=2e-6)
            want = orc.curve_loglik(ref, t, t, vals[m], uncs[m], 1.0, scale_shift=shift)
            assert abs(res.per_meas[k, m, 0] - want) <= 2e-6 * abs(want), (k, m, res.per_meas[k, m, 0], want)
    sf_d = dict(sf, ini_mode="density", _init_params=[np.full(64, 1e16)] * 3)
    with pytest.raises(ValueError, match="fluence"):
        eval_trial_moves(states, np.ones(2), {"TRPL": 1.0}, sf_d)


@pytest.mark.parametrize("tag", ["bounds", "notemper"])
def test_metro_on_gpu_equals_the_reference_golden_chains(tag):
    """metro() on the CUDA evaluator against the golden chains of the unmodified reference's
    metro(serial_fallback=True) (tests/golden/chains.npz): every accept and swap decision, the
    states to rounding, the generator state at the end."""
    from tests import test_golden_chains as gc
    ms = gc.run_ours(tag, factory=None, reference_swap_aliasing=True)
    G = gc.G
    np.testing.assert_array_equal(ms.H.accept, G[f"{tag}_accept"])
    np.testing.assert_allclose(ms.H.states, G[f"{tag}_states"], rtol=gc.STATE_RTOL, atol=0)
    np.testing.assert_array_equal(ms.H.swap_accept, G[f"{tag}_swap_accept"])
    np.testing.assert_allclose(ms.H.loglikelihood, G[f"{tag}_logll"], rtol=5e-3, atol=2e-3)
    rng = np.random.default_rng(0)
    rng.bit_generator.state = ms.random_state
    assert gc.pcg_words(rng) == [int(x) for x in G[f"{tag}_final_rng"]]


def test_single_states_through_the_low_latency_kernel():
    """PathCache(kernel="seulex"): eval_trial_moves / eval_trial_move on one and a few states through
    the extrapolation integrator equal the default kernel's to integration accuracy."""
    import bench
    from metrotrpl_b200.trial_move_evaluation import PathCache, eval_trial_moves
    from tests import parity_cases as pc
    g, prob, params, aux = pc.staub_problem()
    t = g["t"]
    sf = bench.make_shared_fields(g["ini"], t, list(g["vals"]), list(g["uncs"]))
    states = g["states"][[0, 1, 3, 5, 15]]
    a = eval_trial_moves(states, np.ones(5), {"TRPL": 1.0}, sf, cache=PathCache(sf), want_curves=True)
    b = eval_trial_moves(states, np.ones(5), {"TRPL": 1.0}, sf, cache=PathCache(sf, kernel="seulex"), want_curves=True)
    np.testing.assert_allclose(b.logll, a.logll, rtol=1e-6)
    np.testing.assert_allclose(b.curves, a.curves, rtol=1e-5)
    assert b.nsteps[..., 0].mean() < 0.5 * a.nsteps[..., 0].mean()
    one, funcs = eval_trial_move(states[0], {"model_uncertainty": {"TRPL": 1.0}, "_T": 1.0}, sf,
                                 cache=PathCache(sf, kernel="seulex"))
    assert abs(one / a.logll[0] - 1) < 1e-6 and abs(sum(f(1.0) for f in funcs) - one) < 1e-9 * abs(one)
    with pytest.raises(ValueError, match="nx = 128"):
        sim = dict(sf["_sim_info"], nx=[64] * 6)
        PathCache(dict(sf, _sim_info=sim, _init_params=[np.ones(64)] * 6), kernel="seulex")


@pytest.mark.parametrize("kernel", ["auto", "seulex"])
def test_single_chain_on_the_real_staub_data_on_gpu(kernel):
    """configs[0] through the CUDA evaluator (both integrators) against the reference's golden chain."""
    from tests import test_golden_chains as gc
    with tempfile.TemporaryDirectory() as tmp:
        sim_info, ini, e_data, MCMC, param_info = gc.real_chain_problem(tmp)
        ms = metro(sim_info, ini, e_data, MCMC, param_info, export_path="out.pik", install_signal_handlers=False,
                   kernel=kernel)
    gc.check_real_chain(ms)
