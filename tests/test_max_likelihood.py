"""Maximum-likelihood driver (MLE/max_likelihood.py): the batched Nelder-Mead must walk SciPy's
path, and mle() must drive it through the likelihood backend (emu on the CPU tier, CUDA on the GPU
tier)."""
import tempfile

import numpy as np
import pytest
from scipy.optimize import minimize

from metrotrpl_b200.max_likelihood import mle, nelder_mead_batched
from tests.test_metropolis_batched import small_problem, emu_factory


def rosen(x):
    x = np.asarray(x)
    return float(np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))


@pytest.mark.parametrize("x0", [[-1.2, 1.0], [0.3, -0.7, 1.9], [0.0, 0.0, 0.0, 0.0]])
def test_batched_nelder_mead_walks_scipys_path(x0):
    ref = minimize(rosen, np.array(x0, dtype=float), method="Nelder-Mead")
    path = []
    got = nelder_mead_batched(lambda X: [rosen(x) for x in X], x0, callback=lambda x, f: path.append(f))
    assert np.array_equal(got["x"], ref.x)
    assert got["fun"] == ref.fun
    assert got["nit"] == ref.nit and got["nfev"] == ref.nfev
    # one batched call per iteration (plus the shrinks): far fewer launches than evaluations
    assert got["nbatch"] <= got["nit"] + 5
    assert all(b <= a for a, b in zip(path, path[1:]))


def test_mle_with_injected_likelihood_recovers_the_optimum():
    """Host plumbing only: a quadratic log-likelihood in log10 of the active parameters."""
    tmp = tempfile.mkdtemp()
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp, n_chains=1)
    MCMC.pop("parallel_tempering")
    MCMC["current_sigma"] = {"TRPL": 0.05}
    names = param_info["names"]
    act = [i for i, n in enumerate(names) if param_info["active"][n]]
    target = np.array([15.8, -10.1, 2.5, 3.1, 0.7])

    def fake(states):
        return -np.sum((np.log10(states[:, act]) - target) ** 2, axis=1)

    ms = mle(e_data, sim_info, param_info, ini, MCMC, "mle.pik", None, evaluator=fake)
    best = np.log10(ms.H.states[0, act, -1])
    assert np.max(np.abs(best - target)) < 1e-3
    assert ms.H.loglikelihood[0, -1] > -1e-6
    assert ms.opt["message"].startswith("Optimization terminated")
    # inactive parameters never move
    inact = [i for i in range(len(names)) if i not in act]
    assert np.all(ms.H.states[0, inact, :] == ms.H.states[0, inact, :1])


def _mle_on_backend(evaluator):
    """Fit (tauN, ks) from a perturbed start on the small two-curve problem.  The optimiser must
    climb monotonically and end at least as high as the parameters the data were simulated with."""
    tmp = tempfile.mkdtemp()
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp, n_chains=1)
    MCMC.pop("parallel_tempering")
    MCMC["current_sigma"] = {"TRPL": 0.05}
    for n in param_info["active"]:
        param_info["active"][n] = int(n in ("tauN", "ks"))
    truth = np.array([param_info["init_guess"][n] for n in param_info["names"]], dtype=float)
    param_info["init_guess"] = dict(param_info["init_guess"], tauN=300.0, ks=8e-11)
    seen = {}

    def spy(states):
        ll = evaluator(states) if evaluator is not None else seen["cuda"](states)
        return ll

    if evaluator is None:
        # CUDA tier: the same evaluator mle() builds by default, kept to score the truth afterwards
        from metrotrpl_b200.trial_move_evaluation import eval_trial_moves
        holder = {}

        def cuda(states):
            return eval_trial_moves(states, np.ones(len(states)), {"TRPL": 0.05}, holder["ef"]).logll
        seen["cuda"] = cuda
    ms = mle(e_data, sim_info, param_info, ini, MCMC, None, None, evaluator=evaluator)
    ll = ms.H.loglikelihood[0, 1:]
    assert ll[-1] >= ll[0] and np.all(np.diff(ll) >= -1e-9)
    assert ms.opt["nit"] > 10
    if evaluator is None:
        holder["ef"] = ms.ensemble_fields
        ll_truth = float(seen["cuda"](truth[None, :])[0])
    else:
        ll_truth = float(evaluator(truth[None, :])[0])
    assert ll[-1] >= ll_truth - 1e-6 * abs(ll_truth)
    return ms


def test_mle_on_the_emulated_kernel():
    from tests.emu import emu
    from metrotrpl_b200 import _capi
    tmp = tempfile.mkdtemp()
    sim_info, ini, e_data, MCMC, param_info = small_problem(tmp, n_chains=1)
    idx = {n: i for i, n in enumerate(param_info["names"])}
    units = np.array([param_info["unit_conversions"].get(n, 1) for n in param_info["names"]], dtype=float)
    prob = _capi.pack_problem(sim_info, ini, e_data[0], e_data[1], e_data[2])

    def evaluator(states):
        params = _capi.pack_params(states, idx, units)
        aux = _capi.default_aux(len(states), 2, [0.05, 0.05])
        ll, st, ns, _ = emu.loglik_batch(prob, params, aux, _capi.make_opts(RTOL=1e-6), False)
        return ll[:, :, 0].sum(axis=1)

    _mle_on_backend(evaluator)


@pytest.mark.gpu
def test_mle_on_the_gpu():
    _mle_on_backend(None)
