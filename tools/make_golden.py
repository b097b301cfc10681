"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) here.

Run in the build container only (the reference does not travel to the GPU box):

    python tools/make_golden.py            # writes tests/golden/*.npz

Every fixture stores the inputs next to the reference's outputs so the tests can
replay the same inputs through the oracle (CPU) and the CUDA path (GPU).
Nothing here is imported by the product or by the tests.
"""
import logging
import os
import sys
import time

import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)

import forward_solver as ref_fs            # noqa: E402
import laplace as ref_lap                  # noqa: E402
import trial_move_evaluation as ref_tme    # noqa: E402
import utils as ref_utils                  # noqa: E402
from sim_utils import Grid                 # noqa: E402

LOGGER = logging.getLogger("golden")

NAMES = "n0 p0 mu_n mu_p ks Cn Cp Sf Sb tauN tauP eps Tm m".split()
UNITS = np.array([1e-21, 1e-21, 1e5, 1e5, 1e12, 1e33, 1e33, 0.01, 0.01, 1, 1, 1, 1, 1.0])
GUESS = np.array([1e8, 3e15, 20, 20, 4.8e-11, 4.4e-29, 4.4e-29, 10, 10, 511, 871, 10, 300, 1.0])
# prior box of Inputs/mcmc0.txt (inactive parameters pinned to the initial guess)
LO = np.array([1e8, 1e14, 1, 1, 1e-11, 1e-29, 1e-29, 1e-4, 1e-4, 1, 1, 10, 300, 1.0])
HI = np.array([1e8, 1e16, 100, 100, 1e-9, 1e-27, 1e-27, 1e4, 1e4, 1500, 3000, 10, 300, 1.0])
IDX = {n: i for i, n in enumerate(NAMES)}
LENGTHS = [311, 2000, 311, 2000, 311, 2000]
NX = 128


def staub_inputs():
    ini = np.loadtxt(os.path.join(REF, "Inputs", "staub_MAPI_threepower_twothick_input.csv"),
                     delimiter=",")
    d = np.loadtxt(os.path.join(REF, "Inputs", "real_staub_aug_corr_renoised.csv"), delimiter=",")
    t = d[:141, 0]
    t = t[t <= 2000]
    return ini, t


def ref_solve(ini_row, length, t, state, rtol, atol, model="std", names_idx=IDX, units=UNITS,
              nx=NX, meas="TRPL", ini_mode="density", hmax=4):
    g = Grid(length, nx, t, hmax)
    return ref_fs.solve(np.array(ini_row), g, np.array(state, dtype=float), names_idx, meas=meas,
                        units=units, solver=("solveivp",), model=model, ini_mode=ini_mode,
                        RTOL=rtol, ATOL=atol)


def make_shared_fields(ini, times, vals, uncs, lengths, nxs, mtypes, rtol, atol, **extra):
    sim = {"lengths": list(lengths), "nx": list(nxs), "meas_types": list(mtypes),
           "num_meas": len(lengths)}
    sf = {"units": UNITS.copy(), "solver": ("solveivp",), "model": "std", "hmax": 4,
          "rtol": rtol, "atol": atol, "_sim_info": sim, "_param_indexes": dict(IDX),
          "_init_params": np.array(ini, dtype=float), "ini_mode": "density",
          "_times": times, "_vals": vals, "_uncs": uncs}
    sf.update(extra)
    return sf


def gen_staub(n_random=16, seed=20261018):
    ini, t = staub_inputs()
    rng = np.random.default_rng(seed)
    states = [GUESS.copy()]
    for _ in range(n_random):
        states.append(10 ** rng.uniform(np.log10(LO), np.log10(HI)))
    states = np.array(states)
    nS = len(states)
    pl_def = np.zeros((nS, 6, len(t)))
    pl_tight = np.zeros_like(pl_def)
    t0 = time.perf_counter()
    for s in range(nS):
        for m in range(6):
            pl_def[s, m] = ref_solve(ini[m], LENGTHS[m], t, states[s], None, None)
            pl_tight[s, m] = ref_solve(ini[m], LENGTHS[m], t, states[s], 1e-10, 1e-14)
        print(f"staub state {s}/{nS} done ({time.perf_counter() - t0:.0f}s)", flush=True)
    # synthetic measurement: tight-tolerance curve of the initial guess, log10, renoised
    noise = 0.02 * rng.standard_normal((6, len(t)))
    vals = np.log10(pl_tight[0]) + noise
    uncs = np.full((6, len(t)), 0.02) * (1 + 0.5 * rng.random((6, len(t))))
    times = [t.copy() for _ in range(6)]
    sigma = 1.0
    logll = np.zeros(nS)
    logll_T = np.zeros((nS, 3))
    temps = np.array([1.0, 2.0, 8.0])
    for s in range(nS):
        sf = make_shared_fields(ini, times, list(vals), list(uncs), LENGTHS, [NX] * 6, ["TRPL"] * 6,
                                None, None)
        uf = {"model_uncertainty": {"TRPL": sigma}, "_T": 1.0}
        ll, funcs = ref_tme.eval_trial_move(states[s].copy(), uf, sf, LOGGER)
        logll[s] = ll
        for k, T in enumerate(temps):
            logll_T[s, k] = sum(f(T) for f in funcs)
        print(f"staub logll {s}: {ll}", flush=True)
    np.savez_compressed(os.path.join(OUT, "staub6.npz"), names=np.array(NAMES), units=UNITS,
                        states=states, ini=ini, t=t, lengths=np.array(LENGTHS, dtype=float),
                        nx=NX, pl_default=pl_def, pl_tight=pl_tight, vals=vals, uncs=uncs,
                        sigma=sigma, logll=logll, temps=temps, logll_T=logll_T)


def gen_rhs_pins(seed=7):
    rng = np.random.default_rng(seed)
    L = 24
    dx = 3.5
    out = {}
    y = np.abs(rng.standard_normal(3 * L + 1)) * 1e-5
    y[2 * L:] = rng.standard_normal(L + 1) * 1e-4
    args = (L, dx, 1e-13, 3e-6, 2e6, 1.5e6, 4.8e1, 4.4e4, 3.3e4, 0.1, 0.2, 511.0, 871.0, 1.81, 300.0)
    out["std_y"] = y
    out["std_args"] = np.array(args, dtype=float)
    out["std_dy"] = ref_fs.dydt_numba(0.0, y, *args)
    out["std_dy_np"] = ref_fs.dydt(0.0, y, *args)
    y4 = np.abs(rng.standard_normal(4 * L + 1)) * 1e-5
    y4[3 * L:] = rng.standard_normal(L + 1) * 1e-4
    targs = args + (1e3, 1e-6, 50.0)
    out["traps_y"] = y4
    out["traps_args"] = np.array(targs, dtype=float)
    out["traps_dy"] = ref_fs.dydt_numba_traps(0.0, y4, *targs)
    # E field, 1-D and 2-D
    N = np.abs(rng.standard_normal((5, L)))
    P = np.abs(rng.standard_normal((5, L)))
    out["ef_N"] = N
    out["ef_P"] = P
    out["ef_1d"] = ref_fs.E_field(N[0], P[0], 0.1, 0.2, 9.0, dx, corner_E=0.5)
    out["ef_2d"] = ref_fs.E_field(N, P, 0.1, 0.2, 9.0, dx)
    out["int_2d"] = ref_fs.integrate_2D(dx, N)
    out["pl_2d"] = ref_fs.calculate_PL(dx, N, P, 4.8e1, 0.1, 0.2)
    out["trts_2d"] = ref_fs.calculate_TRTS(dx, N, P, 2e6, 1.5e6, 0.1, 0.2)
    np.savez_compressed(os.path.join(OUT, "rhs_pins.npz"), **out)


def gen_known_answers():
    """Replay Tests/test_eval_trial_move.py through the reference and keep its numbers."""
    names = ["n0", "p0", "mu_n", "mu_p", "ks", "Cn", "Cp", "Tm", "Sf", "Sb", "tauN", "tauP",
             "eps", "m"]
    uc = {"n0": 1e-21, "p0": 1e-21, "mu_n": 1e5, "mu_p": 1e5, "ks": 1e12, "Sf": 1e-2, "Sb": 1e-2}
    units = np.array([uc.get(n, 1) for n in names], dtype=float)
    idx = {n: i for i, n in enumerate(names)}
    out = {"names": np.array(names), "units": units}

    def run(guess, lengths, nxs, mtypes, ini, times, vals, uncs, sigma, **extra):
        sim = {"lengths": lengths, "nx": nxs, "meas_types": mtypes, "num_meas": len(lengths)}
        sf = {"units": units, "solver": ("solveivp",), "model": "std", "hmax": 4, "rtol": 1e-5,
              "atol": 1e-8, "_sim_info": sim, "_param_indexes": idx, "_init_params": ini,
              "ini_mode": "density", "_times": times, "_vals": vals, "_uncs": uncs}
        sf.update(extra)
        state = [guess[n] for n in names]
        ll, _ = ref_tme.eval_trial_move(state, {"model_uncertainty": sigma}, sf, LOGGER)
        return ll

    base = {"n0": 0, "p0": 0, "mu_n": 0, "mu_p": 0, "ks": 1e-11, "Sf": 0, "Sb": 0, "Cn": 0,
            "Cp": 0, "Tm": 300, "tauN": 1e99, "tauP": 1e99, "eps": 10, "m": 1}
    nt = 1000
    t100 = np.linspace(0, 100, nt + 1)
    ini2 = np.array([1e15 * np.ones(128), 1e16 * np.ones(128)])
    out["run_iter"] = run(base, [2000, 2000], [128, 128], ["TRPL", "TRPL"], ini2,
                          [t100, t100], [np.ones(nt + 1) * 23] * 2, [np.ones(nt + 1) * 1e-99] * 2,
                          {"TRPL": 1})
    t50 = np.linspace(0, 50, 501)
    out["run_iter_cutoff"] = run(base, [2000, 2000], [128, 128], ["TRPL", "TRPL"], ini2,
                                 [t50, t50], [np.ones(501) * 23] * 2, [np.ones(501) * 1e-99] * 2,
                                 {"TRPL": 1})
    dep = dict(base, n0=1e8, p0=1e17, ks=1e-13, tauN=4, tauP=4)
    valsd = [np.log10(2e14 * np.exp(-t100 / 8))]
    ini1 = np.array([1e15 * np.ones(128)])
    out["depletion_4"] = run(dep, [2000], [128], ["TRPL"], ini1, [t100], valsd,
                             [np.ones(nt + 1) * 1e-99], {"TRPL": 1}, force_min_y=True)
    dep2 = dict(dep, tauN=4.01, tauP=4.01)
    out["depletion_401"] = run(dep2, [2000], [128], ["TRPL"], ini1, [t100], valsd,
                               [np.ones(nt + 1) * 1e-99], {"TRPL": 1}, force_min_y=True)
    mixed = dict(base, mu_n=0.01, mu_p=0.01)
    ini3 = np.array([1e15 * np.ones(128), 1e15 * np.ones(128)])
    out["mixed_types"] = run(mixed, [2000, 2000], [128, 128], ["TRPL", "TRTS"], ini3,
                             [t100, t100], [np.ones(nt + 1) * 23, np.ones(nt + 1) * -2],
                             [np.ones(nt + 1) * 1e-99] * 2, {"TRPL": 1, "TRTS": 10})
    # raw curves of the first case for curve-level parity (rtol 1e-5 and tight)
    st = np.array([base[n] for n in names], dtype=float)
    g = Grid(2000, 128, t100, 4)
    out["run_iter_pl0"] = ref_fs.solve(ini2[0].copy(), g, st.copy(), idx, units=units, RTOL=1e-5, ATOL=1e-8)
    out["run_iter_pl0_tight"] = ref_fs.solve(ini2[0].copy(), g, st.copy(), idx, units=units,
                                             RTOL=1e-11, ATOL=1e-16)
    st = np.array([mixed[n] for n in names], dtype=float)
    out["mixed_trts_tight"] = ref_fs.solve(ini3[1].copy(), g, st.copy(), idx, meas="TRTS",
                                           units=units, RTOL=1e-11, ATOL=1e-16)
    st = np.array([dep[n] for n in names], dtype=float)
    out["depletion_pl_tight"] = ref_fs.solve(ini1[0].copy(), g, st.copy(), idx, units=units,
                                             RTOL=1e-11, ATOL=1e-16)
    np.savez_compressed(os.path.join(OUT, "known_answers.npz"), **out)


def gen_irf_pins():
    irf = np.loadtxt(os.path.join(REF, "IRFs", "irf_520nm.csv"), delimiter=",")
    tables = ref_lap.make_I_tables({520: irf})
    mom, t_irf = tables[520]
    t = np.linspace(0, 40, 161)
    y = 3e18 * (0.7 * np.exp(-t / 1.5) + 0.3 * np.exp(-t / 25.0))
    ct, cy, ok = ref_lap.do_irf_convolution(t, y, tables[520], time_max_shift=True)
    exp_t = np.linspace(0, 38, 153)
    exp_y = np.log10(y[:153]) + 0.01
    exp_u = np.full(153, 0.03)
    sy, tc, vc, uc = ref_lap.post_conv_trim(ct, cy, exp_t, exp_y, exp_u)
    sol = np.array([5.0, 4.0, 3.0, 2.0, 1.5, 1.2, 0.5, 0.1])
    vals = np.log10(np.array([5.0, 4.0, 3.0, 2.0, 1.5, 1.3, 1.25, 1.21]))
    s2, floor, nset = ref_utils.set_min_y(sol.copy(), vals, 0.1)
    np.savez_compressed(os.path.join(OUT, "irf_pins.npz"), irf=irf, moments=mom, t_irf=t_irf,
                        t=t, y=y, conv_t=ct, conv_y=cy, conv_ok=ok, exp_t=exp_t, exp_y=exp_y,
                        exp_u=exp_u, trim_y=sy, trim_t=tc, trim_v=vc, trim_u=uc,
                        minY_sol=sol, minY_vals=vals, minY_out=s2, minY_floor=floor,
                        minY_nset=nset)


def gen_traps_irf(seed=11):
    """configs[3]: trap-assisted model + IRF convolution, nx=256, stiff capture."""
    names = NAMES[:-1] + ["kC", "Nt", "tauE"]
    units = np.concatenate([UNITS[:-1], [1e12, 1e-21, 1.0]])
    idx = {n: i for i, n in enumerate(names)}
    irf = np.loadtxt(os.path.join(REF, "IRFs", "irf_520nm.csv"), delimiter=",")
    tables = ref_lap.make_I_tables({520: irf})
    nx = 256
    t = np.linspace(0, 100, 401)
    base = np.concatenate([GUESS[:-1], [1e-7, 3e15, 200.0]])
    rng = np.random.default_rng(seed)
    states = [base]
    for _ in range(3):
        s = base.copy()
        s[1:11] = 10 ** rng.uniform(np.log10(LO[1:11]), np.log10(HI[1:11]))
        s[-3:] = [10 ** rng.uniform(-9, -6), 10 ** rng.uniform(14, 16.5), 10 ** rng.uniform(0.5, 3)]
        states.append(s)
    states = np.array(states)
    inis = np.array([[2e12, 6e4, 1], [3e13, 6e4, 1]], dtype=float)  # fluence, alpha, direction
    lengths = [311.0, 2000.0]
    pl_def = np.zeros((len(states), 2, len(t)))
    pl_tight = np.zeros_like(pl_def)
    for s in range(len(states)):
        for m in range(2):
            pl_def[s, m] = ref_solve(inis[m], lengths[m], t, states[s], None, None, model="traps",
                                     names_idx=idx, units=units, nx=nx, ini_mode="fluence")
            pl_tight[s, m] = ref_solve(inis[m], lengths[m], t, states[s], 1e-10, 1e-15,
                                       model="traps", names_idx=idx, units=units, nx=nx,
                                       ini_mode="fluence")
        print("traps state", s, flush=True)
    # measurement = convolved + trimmed tight curve of state 0 (same grid as the simulation,
    # as in the reference where shared_fields["_times"] is both), log10, renoised
    vals = []
    uncs = []
    for m in range(2):
        ct, cy, ok = ref_lap.do_irf_convolution(t, pl_tight[0, m].copy(), tables[520],
                                                time_max_shift=True)
        sy, tc, _, _ = ref_lap.post_conv_trim(ct, cy, t, t * 0, t * 0)
        v = np.full(len(t), np.log10(np.abs(sy[-1])))
        v[:len(sy)] = np.log10(np.abs(sy))
        vals.append(v + 0.02 * rng.standard_normal(len(t)))
        uncs.append(np.full(len(t), 0.03))
    logll = np.zeros(len(states))
    for s in range(len(states)):
        sim = {"lengths": lengths, "nx": [nx, nx], "meas_types": ["TRPL", "TRPL"], "num_meas": 2}
        sf = {"units": units, "solver": ("solveivp",), "model": "traps", "hmax": 4, "rtol": None,
              "atol": None, "_sim_info": sim, "_param_indexes": idx,
              "_init_params": inis.copy(), "ini_mode": "fluence", "_times": [t, t],
              "_vals": [v.copy() for v in vals], "_uncs": [u.copy() for u in uncs],
              "irf_convolution": [520, 520], "_IRF_tables": tables}
        ll, _ = ref_tme.eval_trial_move(states[s].copy(), {"model_uncertainty": {"TRPL": 1.0}},
                                        sf, LOGGER)
        logll[s] = ll
        print("traps logll", s, ll, flush=True)
    np.savez_compressed(os.path.join(OUT, "traps_irf.npz"), names=np.array(names), units=units,
                        states=states, inis=inis, lengths=np.array(lengths), nx=nx, t=t, irf=irf,
                        moments=tables[520][0], t_irf=tables[520][1], pl_default=pl_def,
                        pl_tight=pl_tight, vals=np.array(vals), uncs=np.array(uncs), logll=logll)


def gen_real3():
    """configs[0] on the reference's real data: Inputs/real_staub_input.csv (three injection levels,
    311 nm film) against Inputs/real_staub_aug_corr_renoised.csv, both read by the reference's own
    bayes_io.get_data / get_initpoints with mcmc0.txt's flags (time cutoff 0..2000 ns, log10 of the
    measurement and its uncertainty), parameter set list = the staub6 fixture's 17 states.  (The
    six-curve measurement file mcmc0.txt names, staub_MAPI_threepower_twothick_withauger.csv, is
    not part of the reference tree; this three-curve set is the real data it ships.)"""
    import bayes_io as ref_io
    g6 = np.load(os.path.join(OUT, "staub6.npz"))
    states = g6["states"]
    ic_flags = {"time_cutoff": [0, 2000], "select_obs_sets": None, "noise_level": None}
    times, vals, uncs = ref_io.get_data(os.path.join(REF, "Inputs", "real_staub_aug_corr_renoised.csv"),
                                        ic_flags, {"log_y": 1})
    ini = ref_io.get_initpoints(os.path.join(REF, "Inputs", "real_staub_input.csv"), ic_flags)
    lengths = [311.0, 311.0, 311.0]
    nS = len(states)
    temps = np.array([1.0, 2.0, 8.0])
    sigma = 1.0
    pl_def = [[None] * 3 for _ in range(nS)]
    logll = np.zeros(nS)
    logll_T = np.zeros((nS, 3))
    for s in range(nS):
        for m in range(3):
            pl_def[s][m] = ref_solve(ini[m], lengths[m], times[m], states[s], None, None)
        sf = make_shared_fields(ini, times, vals, uncs, lengths, [NX] * 3, ["TRPL"] * 3, None, None)
        ll, funcs = ref_tme.eval_trial_move(states[s].copy(), {"model_uncertainty": {"TRPL": sigma}, "_T": 1.0},
                                            sf, LOGGER)
        logll[s] = ll
        logll_T[s] = [sum(f(T) for f in funcs) for T in temps]
        print("real3 state", s, ll, flush=True)
    n_t = np.array([len(tt) for tt in times])

    def pad(rows):
        return np.array([np.pad(np.asarray(r, dtype=float), (0, n_t.max() - len(r)), constant_values=np.nan)
                         for r in rows])
    np.savez_compressed(os.path.join(OUT, "staub_real3.npz"), names=np.array(NAMES), units=UNITS, states=states,
                        ini=ini, lengths=np.array(lengths), nx=NX, n_t=n_t, t=pad(times), vals=pad(vals),
                        uncs=pad(uncs), sigma=sigma, temps=temps, logll=logll, logll_T=logll_T,
                        pl_default=np.array([pad(pl_def[s]) for s in range(nS)]))


def gen_chains():
    """Golden chains: the reference's own metro(serial_fallback=True) (metropolis.py:283-473 ->
    main_metro_loop_serial :93-137, trial_displacement_move :42-63, swap_move_serial :66-90,
    trial_move_generation.make_trial_move :54-96), unmodified, on the small problem of
    tests/test_metropolis_batched.small_problem (nx = 32, two curves, 4 chains, 20 iterations).
    Recorded by wrapping (not replacing) the reference's functions: every proposal with the
    generator state before it, every acceptance draw with its log-ratio, every swap attempt, the
    final History.  Variants: hard bounds on (narrow prior: many retries), hard bounds off,
    do_mu_constraint on (global np.random seeded), tempering off."""
    import copy
    import tempfile
    sys.path.insert(0, os.path.join(HERE, ".."))
    import metropolis as ref_metro
    from tests.test_metropolis_batched import small_problem

    def pcg_words(state):
        s = state["state"]
        m = (1 << 64) - 1
        return np.array([s["state"] >> 64, s["state"] & m, s["inc"] >> 64, s["inc"] & m,
                         state["has_uint32"], state["uinteger"]], dtype=np.uint64)

    def replay_u(state):
        g = np.random.Generator(np.random.PCG64())
        g.bit_generator.state = state
        return g.random()

    out = {}
    variants = {
        "bounds": dict(hard_bounds=1, narrow=True, temper=True, mu=None),
        "free": dict(hard_bounds=0, narrow=True, temper=True, mu=None),
        "mu": dict(hard_bounds=1, narrow=False, temper=False, mu=(20.0, 3.0)),
        "notemper": dict(hard_bounds=1, narrow=True, temper=False, mu=None),
    }
    for tag, v in variants.items():
        tmp = tempfile.mkdtemp()
        sim_info, ini, e_data, MCMC, param_info = small_problem(tmp, n_chains=4, num_iters=20)
        MCMC["checkpoint_freq"] = 20
        MCMC["hard_bounds"] = v["hard_bounds"]
        if v["narrow"]:
            param_info["prior_dist"]["p0"] = (2e15, 4e15)
            param_info["prior_dist"]["tauN"] = (400, 650)
            param_info["prior_dist"]["Sf"] = (5, 20)
        if not v["temper"]:
            MCMC["temper_freq"] = 1000
        if v["mu"] is not None:
            MCMC["do_mu_constraint"] = v["mu"]
            param_info["active"]["mu_n"] = 1
            param_info["active"]["mu_p"] = 1
        rec = {"prop_in": [], "prop_move": [], "prop_out": [], "prop_rng": [], "u": [], "logratio": [],
               "acc": [], "acc_rng": [], "swap_i": [], "swap_k": [], "swap_ok": []}
        orig_mtm, orig_roll, orig_swap = ref_metro.make_trial_move, ref_metro.roll_acceptance, ref_metro.swap_move_serial

        def mtm(cur, move, sf, RNG, logger):
            rec["prop_rng"].append(pcg_words(RNG.bit_generator.state))
            rec["prop_in"].append(np.array(cur, dtype=float))
            rec["prop_move"].append(np.array(move, dtype=float))
            new = orig_mtm(cur, move, sf, RNG, logger)
            rec["prop_out"].append(np.array(new, dtype=float))
            return new

        def roll(rng, logratio):
            st = copy.deepcopy(rng.bit_generator.state)
            res = orig_roll(rng, logratio)
            rec["acc_rng"].append(pcg_words(st))
            rec["u"].append(replay_u(st))
            rec["logratio"].append(float(logratio))
            rec["acc"].append(bool(res))
            return res

        def swap(k, i, *a, **kw):
            res = orig_swap(k, i, *a, **kw)
            rec["swap_i"].append(int(i)); rec["swap_k"].append(int(k)); rec["swap_ok"].append(bool(res))
            return res
        ref_metro.make_trial_move, ref_metro.roll_acceptance, ref_metro.swap_move_serial = mtm, roll, swap
        ref_metro.all_signal_handler = lambda f: None          # leave this process's signal handlers alone
        np.random.seed(1234)                                   # the reference's mu constraint draws from np.random
        try:
            ref_metro.metro(sim_info, ini, e_data, copy.deepcopy(MCMC), copy.deepcopy(param_info),
                            export_path="gold.pik", serial_fallback=True)
        finally:
            ref_metro.make_trial_move, ref_metro.roll_acceptance, ref_metro.swap_move_serial = orig_mtm, orig_roll, orig_swap
        import pickle
        with open(os.path.join(tmp, "gold.pik"), "rb") as f:
            sys.path.insert(0, REF)
            MS = pickle.load(f)
        for k, val in rec.items():
            out[f"{tag}_{k}"] = np.array(val)
        out[f"{tag}_states"] = MS.H.states
        out[f"{tag}_logll"] = MS.H.loglikelihood
        out[f"{tag}_accept"] = MS.H.accept
        out[f"{tag}_swap_accept"] = MS.H.swap_accept
        out[f"{tag}_swap_attempts"] = MS.H.swap_attempts
        out[f"{tag}_final_rng"] = pcg_words(MS.random_state)
        out[f"{tag}_vals"] = np.array(e_data[1])
        print(tag, "accept", MS.H.accept.sum(axis=1), "swaps", MS.H.swap_accept, "proposals", len(rec["prop_out"]),
              "acceptance draws", len(rec["u"]), flush=True)
    np.savez_compressed(os.path.join(OUT, "chains.npz"), **out)


def gen_chain_real():
    """configs[0]: a single chain of the reference's metro(serial_fallback=True) on the REAL staub
    measurement (tests/golden/staub_real3.npz: Inputs/real_staub_input.csv against
    Inputs/real_staub_aug_corr_renoised.csv, nx = 128, three curves) with Inputs/mcmc0.txt's parameter
    names, unit conversions, log-scale and activity flags, priors and initial guess; box half-width
    0.1 decades, model uncertainty 1, hard bounds, 16 iterations."""
    import copy
    import pickle
    import tempfile
    import metropolis as ref_metro
    g = np.load(os.path.join(OUT, "staub_real3.npz"))
    n_t = g["n_t"]
    e_data = ([g["t"][m, :n_t[m]] for m in range(3)], [g["vals"][m, :n_t[m]] for m in range(3)],
              [g["uncs"][m, :n_t[m]] for m in range(3)])
    sim_info = {"lengths": [311.0] * 3, "nx": [NX] * 3, "meas_types": ["TRPL"] * 3, "num_meas": 3}
    active = {n: int(n not in ("n0", "eps", "Tm", "m")) for n in NAMES}            # mcmc0.txt "Active"
    param_info = {"names": list(NAMES), "active": active, "unit_conversions": dict(zip(NAMES, UNITS)),
                  "do_log": {n: 1 for n in NAMES},
                  "prior_dist": {n: ((lo, hi) if active[n] else (0, np.inf)) for n, lo, hi in zip(NAMES, LO, HI)},
                  "init_guess": dict(zip(NAMES, GUESS)), "trial_move": {n: 0.1 for n in NAMES}}
    param_info["prior_dist"]["m"] = (-np.inf, np.inf)
    tmp = tempfile.mkdtemp()
    MCMC = {"init_cond_path": "real_staub_input.csv", "measurement_path": "real_staub_aug_corr_renoised.csv",
            "output_path": tmp, "num_iters": 16, "solver": ("solveivp",), "model": "std", "ini_mode": "density",
            "log_y": 1, "checkpoint_freq": 16, "hard_bounds": 1, "rtol": None, "atol": None,
            "model_uncertainty": {"TRPL": 1.0}}
    ref_metro.all_signal_handler = lambda f: None
    ref_metro.metro(copy.deepcopy(sim_info), g["ini"].copy(), e_data, copy.deepcopy(MCMC), copy.deepcopy(param_info),
                    export_path="gold.pik", serial_fallback=True)
    with open(os.path.join(tmp, "gold.pik"), "rb") as f:
        MS = pickle.load(f)
    print("real-data chain: accepted", int(MS.H.accept.sum()), "of 15; logll", MS.H.loglikelihood[0])
    np.savez_compressed(os.path.join(OUT, "chain_real.npz"), states=MS.H.states, logll=MS.H.loglikelihood,
                        accept=MS.H.accept, num_iters=16, trial_move=0.1, sigma=1.0)


if __name__ == "__main__":
    which = sys.argv[1:] or ["rhs", "irf", "known", "staub", "traps", "real3", "chains", "chain_real"]
    if "rhs" in which:
        gen_rhs_pins()
    if "irf" in which:
        gen_irf_pins()
    if "known" in which:
        gen_known_answers()
    if "staub" in which:
        gen_staub()
    if "traps" in which:
        gen_traps_irf()
    if "real3" in which:
        gen_real3()
    if "chains" in which:
        gen_chains()
    if "chain_real" in which:
        gen_chain_real()
