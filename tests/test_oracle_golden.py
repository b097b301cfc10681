"""Pin the CPU oracle against fixtures produced by the real reference (tools/make_golden.py)
and against the known answers in the reference's own unit tests."""
import os

import numpy as np
import pytest

from oracle import trpl_oracle as orc

pytestmark = pytest.mark.filterwarnings("ignore")


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_rhs_std_bit_exact(golden_dir):
    g = load(golden_dir, "rhs_pins.npz")
    a = g["std_args"]
    dy = orc.rhs_std(0.0, g["std_y"], int(a[0]), *a[1:])
    np.testing.assert_array_equal(dy, g["std_dy"])
    # the reference's own numpy twin agrees with its numba version to rounding only
    np.testing.assert_allclose(dy, g["std_dy_np"], rtol=1e-12, atol=1e-30)


def test_rhs_traps_bit_exact(golden_dir):
    g = load(golden_dir, "rhs_pins.npz")
    a = g["traps_args"]
    dy = orc.rhs_traps(0.0, g["traps_y"], int(a[0]), *a[1:])
    np.testing.assert_array_equal(dy, g["traps_dy"])


def test_e_field_and_integrals(golden_dir):
    g = load(golden_dir, "rhs_pins.npz")
    N, P = g["ef_N"], g["ef_P"]
    np.testing.assert_array_equal(orc.e_field(N[0], P[0], 0.1, 0.2, 9.0, 3.5, corner_E=0.5), g["ef_1d"])
    np.testing.assert_array_equal(orc.e_field(N, P, 0.1, 0.2, 9.0, 3.5), g["ef_2d"])
    np.testing.assert_array_equal(orc.integrate_nodes(3.5, N), g["int_2d"])
    np.testing.assert_array_equal(orc.pl_signal(3.5, N, P, 4.8e1, 0.1, 0.2), g["pl_2d"])
    np.testing.assert_array_equal(orc.trts_signal(3.5, N, P, 2e6, 1.5e6, 0.1, 0.2), g["trts_2d"])
    with pytest.raises(NotImplementedError):
        orc.e_field(np.zeros((2, 2, 2)), np.zeros((2, 2, 2)), 0, 0, 1, 1)


def test_e_field_reference_unit_cases():
    # Tests/test_metropolis.py:37-90
    nx, dx = 10, 1
    q_over_eps = orc.Q_C / orc.EPS0
    N = np.zeros(nx)
    np.testing.assert_equal(orc.e_field(N, N, 0, 0, 1, dx), np.zeros(nx + 1))
    N = np.ones(nx)
    E = orc.e_field(N, 2 * N, 0, 0, 1, dx)
    np.testing.assert_equal(E[1:], q_over_eps * np.cumsum(N))
    E = orc.e_field(N, -N, 0, 0, 1, dx, corner_E=24)
    np.testing.assert_equal(E[1:], -2 * q_over_eps * np.cumsum(N) + 24)


def test_irf_tables_and_convolution(golden_dir):
    g = load(golden_dir, "irf_pins.npz")
    tab = orc.irf_moment_tables({520: g["irf"]})[520]
    np.testing.assert_array_equal(tab[0], g["moments"])
    ct, cy, ok = orc.irf_convolve(g["t"], g["y"], (g["moments"], g["t_irf"]), time_max_shift=True)
    assert ok == bool(g["conv_ok"])
    np.testing.assert_array_equal(ct, g["conv_t"])
    np.testing.assert_allclose(cy, g["conv_y"], rtol=1e-13)
    sy, tc, vc, uc = orc.trim_after_convolution(g["conv_t"], g["conv_y"], g["exp_t"], g["exp_y"], g["exp_u"])
    np.testing.assert_array_equal(tc, g["trim_t"])
    np.testing.assert_allclose(sy, g["trim_y"], rtol=1e-13)
    np.testing.assert_array_equal(vc, g["trim_v"])
    np.testing.assert_array_equal(uc, g["trim_u"])


def test_convolution_analytic():
    # Tests/test_convolution.py test1: exp(-t) * sin(t) = (exp(-t) + sin t - cos t)/2
    t = np.linspace(0, 10, 1001)
    dt = t[1] - t[0]
    tt = np.arange(0, t[-1] + dt / 4, dt / 2)
    g_t = np.sin(t)
    mom = np.zeros((len(t), 3))
    for i in range(len(t) - 1):
        for n in range(3):
            mom[i, n] = orc.irf_moment(t, g_t, i, n, u_spacing=1000)
    h = orc.moment_convolve(np.exp(-tt), mom)
    np.testing.assert_almost_equal(h, 0.5 * (np.exp(-t) + np.sin(t) - np.cos(t)), decimal=5)


def test_min_y(golden_dir):
    g = load(golden_dir, "irf_pins.npz")
    s, floor, n = orc.raise_to_min_y(g["minY_sol"].copy(), g["minY_vals"], 0.1)
    np.testing.assert_array_equal(s, g["minY_out"])
    assert floor == float(g["minY_floor"]) and n == int(g["minY_nset"])


def _known_setup(g):
    names = [str(n) for n in g["names"]]
    return names, g["units"], {n: i for i, n in enumerate(names)}


def test_known_answers_run_iter(golden_dir):
    """Tests/test_eval_trial_move.py:21-80 (expects -59340.105083 + -32560.139058, decimal=0)."""
    g = load(golden_dir, "known_answers.npz")
    names, units, idx = _known_setup(g)
    base = {"n0": 0, "p0": 0, "mu_n": 0, "mu_p": 0, "ks": 1e-11, "Sf": 0, "Sb": 0, "Cn": 0,
            "Cp": 0, "Tm": 300, "tauN": 1e99, "tauP": 1e99, "eps": 10, "m": 1}
    state = np.array([base[n] for n in names], dtype=float)
    t = np.linspace(0, 100, 1001)
    sim = {"lengths": [2000, 2000], "nx": [128, 128], "meas_types": ["TRPL", "TRPL"], "num_meas": 2}
    ini = np.array([1e15 * np.ones(128), 1e16 * np.ones(128)])
    ll, per = orc.state_loglik(state, sim, ini, [t, t], [np.ones(1001) * 23] * 2,
                               [np.ones(1001) * 1e-99] * 2, idx, units, {"TRPL": 1.0},
                               rtol=1e-5, atol=1e-8)
    np.testing.assert_almost_equal(ll, np.sum([-59340.105083, -32560.139058]), decimal=0)
    np.testing.assert_allclose(ll, float(g["run_iter"]), rtol=1e-12)
    # curve-level: same LSODA, same inputs -> identical curve
    gr = orc.Grid(2000, 128, t, 4)
    pl = orc.simulate(ini[0], gr, state, idx, units=units, RTOL=1e-5, ATOL=1e-8)
    np.testing.assert_allclose(pl, g["run_iter_pl0"], rtol=1e-12)


def test_known_answers_depletion_and_mixed(golden_dir):
    g = load(golden_dir, "known_answers.npz")
    names, units, idx = _known_setup(g)
    t = np.linspace(0, 100, 1001)
    dep = {"n0": 1e8, "p0": 1e17, "mu_n": 0, "mu_p": 0, "ks": 1e-13, "Sf": 0, "Sb": 0, "Cn": 0,
           "Cp": 0, "Tm": 300, "tauN": 4, "tauP": 4, "eps": 10, "m": 1}
    sim = {"lengths": [2000], "nx": [128], "meas_types": ["TRPL"], "num_meas": 1}
    ini = np.array([1e15 * np.ones(128)])
    vals = [np.log10(2e14 * np.exp(-t / 8))]
    lls = []
    for tau, key in ((4, "depletion_4"), (4.01, "depletion_401")):
        st = dict(dep, tauN=tau, tauP=tau)
        state = np.array([st[n] for n in names], dtype=float)
        ll, _ = orc.state_loglik(state, sim, ini, [t], vals, [np.ones(1001) * 1e-99], idx, units,
                                 {"TRPL": 1.0}, rtol=1e-5, atol=1e-8, force_min_y=True)
        np.testing.assert_allclose(ll, float(g[key]), rtol=1e-10)
        lls.append(ll)
    assert lls[1] > lls[0]          # Tests/test_eval_trial_move.py:143
    mixed = {"n0": 0, "p0": 0, "mu_n": 0.01, "mu_p": 0.01, "ks": 1e-11, "Sf": 0, "Sb": 0, "Cn": 0,
             "Cp": 0, "Tm": 300, "tauN": 1e99, "tauP": 1e99, "eps": 10, "m": 1}
    state = np.array([mixed[n] for n in names], dtype=float)
    sim = {"lengths": [2000, 2000], "nx": [128, 128], "meas_types": ["TRPL", "TRTS"], "num_meas": 2}
    ini = np.array([1e15 * np.ones(128), 1e15 * np.ones(128)])
    ll, _ = orc.state_loglik(state, sim, ini, [t, t], [np.ones(1001) * 23, np.ones(1001) * -2],
                             [np.ones(1001) * 1e-99] * 2, idx, units, {"TRPL": 1.0, "TRTS": 10.0},
                             rtol=1e-5, atol=1e-8)
    np.testing.assert_almost_equal(ll, np.sum([-59340.105083, -517.98]), decimal=0)
    np.testing.assert_allclose(ll, float(g["mixed_types"]), rtol=1e-10)


def test_staub_curves_and_loglik(golden_dir):
    """Same LSODA + same RHS on the example's 6 curves must reproduce the reference."""
    path = os.path.join(golden_dir, "staub6.npz")
    if not os.path.exists(path):
        pytest.skip("staub6 fixture not generated")
    g = np.load(path)
    names = [str(n) for n in g["names"]]
    idx = {n: i for i, n in enumerate(names)}
    t = g["t"]
    for s in (0, 3):
        for m in (0, 5):
            gr = orc.Grid(g["lengths"][m], int(g["nx"]), t, 4)
            pl = orc.simulate(g["ini"][m], gr, g["states"][s], idx, units=g["units"])
            # LSODA's dense LU goes through threaded LAPACK: its step sequence, hence its output,
            # is reproducible only to about its own tolerance across BLAS thread counts
            np.testing.assert_allclose(pl, g["pl_default"][s, m], rtol=5e-6)
    sim = {"lengths": list(g["lengths"]), "nx": [int(g["nx"])] * 6, "meas_types": ["TRPL"] * 6,
           "num_meas": 6}
    ll, _ = orc.state_loglik(g["states"][0], sim, g["ini"], [t] * 6, list(g["vals"]), list(g["uncs"]),
                             idx, g["units"], {"TRPL": float(g["sigma"])})
    np.testing.assert_allclose(ll, g["logll"][0], rtol=5e-6)
