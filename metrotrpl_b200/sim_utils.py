"""Grid, History, Ensemble and the constants of the reference's sim_utils.py that the path touches.

When this package is dropped into a MetroTRPL checkout the reference's own
`sim_utils` is importable and ITS History / Ensemble / Grid are used - existing checkpoints then
unpickle into the classes that wrote them and nothing is duplicated.  Stand-alone (tests, bench,
the GPU box) the minimal twins below are used: /root/reference/sim_utils.py:13-23 (constants),
:25-99 (History), :100-196 (Ensemble), :248-283 (Grid).
"""
import pickle
from sys import float_info

import numpy as np

DEFAULT_HMAX = 4            # sim_utils.py:17
DEFAULT_TEMPER_FREQ = 10    # sim_utils.py:19
MAX_PROPOSALS = 100         # sim_utils.py:20
NEGATIVE_FRAC_TOL = 0.2     # sim_utils.py:23


class Grid:
    """Space and time grid of one measurement (same attributes as the reference's Grid)."""

    def __init__(self, thickness, nx, tSteps, hmax=DEFAULT_HMAX):
        self.thickness = thickness
        self.nx = nx
        self.dx = self.thickness / self.nx
        self.xSteps = np.linspace(self.dx / 2, self.thickness - self.dx / 2, self.nx)
        if tSteps[0] != 0:
            raise ValueError("Grid error - times must start at t=0")
        self.tSteps = tSteps
        self.start_time = 0
        self.nt = len(tSteps) - 1
        self.hmax = hmax
        self.final_time = self.tSteps[-1]
        self.min_y = float_info.min


class History:
    """Visited states, acceptance flags and log-likelihoods of every chain (sim_utils.py:25-75)."""

    def __init__(self, n_chains, num_iters, names):
        self.states_are_one_array = True
        self.states = np.zeros((n_chains, len(names), num_iters), dtype=float)
        self.accept = np.zeros((n_chains, num_iters), dtype=int)
        self.loglikelihood = np.zeros((n_chains, num_iters), dtype=float)
        self.swap_attempts = np.zeros(n_chains, dtype=int)
        self.swap_accept = np.zeros(n_chains, dtype=int)

    def update(self, names):
        for i, param in enumerate(names):
            setattr(self, f"mean_{param}", self.states[:, i])

    def pack(self, states, logll, accept):
        self.states = states
        self.loglikelihood = logll
        self.accept = accept

    def truncate(self, k):
        self.states = self.states[:, :, :k]
        self.accept = self.accept[:, :k]
        self.loglikelihood = self.loglikelihood[:, :k]

    def extend(self, new_num_iters):
        cur = len(self.accept[0])
        if new_num_iters < cur:
            self.truncate(new_num_iters)
            return
        if new_num_iters == cur:
            return
        extra = new_num_iters - cur
        self.accept = np.concatenate((self.accept, np.zeros((self.accept.shape[0], extra))), axis=1)
        self.loglikelihood = np.concatenate(
            (self.loglikelihood, np.zeros((self.loglikelihood.shape[0], extra))), axis=1)
        self.states = np.concatenate(
            (self.states, np.zeros((self.states.shape[0], self.states.shape[1], extra))), axis=2)


class Ensemble:
    """Chains of a (parallel-tempering) run and the fields they share (sim_utils.py:77-210).

    Consumes param_info / MCMC_fields exactly as the reference does (the dicts are popped).
    """

    def __init__(self, param_info, sim_info, MCMC_fields, num_iters, verbose=False):
        ef = {}
        for f in ["output_path", "init_cond_path", "measurement_path", "checkpoint_freq", "ini_mode",
                  "solver", "model", "num_iters", "log_y"]:
            ef[f] = MCMC_fields.pop(f)
        for f in ["rtol", "atol", "scale_factor", "load_checkpoint", "fittable_fluences",
                  "fittable_absps", "irf_convolution", "do_mu_constraint"]:
            ef[f] = MCMC_fields.pop(f, None)
        ef["temper_freq"] = MCMC_fields.pop("temper_freq", DEFAULT_TEMPER_FREQ)
        if "model_uncertainty" in MCMC_fields and "likel2move_ratio" in MCMC_fields:
            MCMC_fields.pop("likel2move_ratio")
        if "likel2move_ratio" in MCMC_fields:
            ef["likel2move_ratio"] = MCMC_fields.pop("likel2move_ratio")
        ef["hard_bounds"] = MCMC_fields.pop("hard_bounds", 0)
        ef["hmax"] = MCMC_fields.pop("hmax", DEFAULT_HMAX)
        ef["force_min_y"] = MCMC_fields.pop("force_min_y", 0)
        names = param_info["names"]
        ef["prior_dist"] = param_info.pop("prior_dist")
        do_log = param_info.pop("do_log")
        ef["do_log"] = np.array([do_log[p] for p in names], dtype=bool)
        trial = param_info.pop("trial_move")
        ef["base_trial_move"] = np.array(
            [trial[p] if param_info["active"][p] else 0 for p in names], dtype=float)
        active = param_info.pop("active")
        ef["active"] = np.array([active[p] for p in names], dtype=bool)
        units = param_info.pop("unit_conversions")
        ef["units"] = np.array([units.get(p, 1) for p in names], dtype=float)
        ef["_param_indexes"] = {name: names.index(name) for name in names}
        ef["_T"] = MCMC_fields.pop("parallel_tempering", [1])
        ef["_n_chains"] = len(ef["_T"])
        ef["names"] = param_info.pop("names")
        init_state = np.array([param_info["init_guess"][p] for p in ef["names"]], dtype=float)
        self.ensemble_fields = ef
        self.H = History(ef["_n_chains"], num_iters, ef["names"])
        self.H.states[:, :, 0] = init_state
        self.unique_fields = []
        for i in range(ef["_n_chains"]):
            uf = dict(MCMC_fields)
            uf["_T"] = ef["_T"][i]
            if "likel2move_ratio" in ef:
                uf["model_uncertainty"] = {m: max(ef["base_trial_move"]) * ef["likel2move_ratio"][m]
                                           for m in sim_info["meas_types"]}
            self.unique_fields.append(uf)
        ef["do_parallel_tempering"] = ef["_n_chains"] > 1
        ef["_sim_info"] = sim_info
        self.latest_iter = 0
        self.random_state = None

    def checkpoint(self, fname):
        """Pickle the ensemble (sim_utils.py:93-99)."""
        self.H.update(self.ensemble_fields["names"])
        with open(fname, "wb+") as f:
            pickle.dump(self, f)


def _bind_reference_classes():
    """Prefer the reference's own classes when its sim_utils is importable (and really is it)."""
    global Grid, History, Ensemble
    try:
        import sim_utils as ref
    except Exception:
        return False
    if ref is globals().get("__module_self__"):
        return False
    need = ("Grid", "History", "Ensemble", "MAX_PROPOSALS", "NEGATIVE_FRAC_TOL", "DEFAULT_HMAX")
    if not all(hasattr(ref, n) for n in need) or ref.MAX_PROPOSALS != MAX_PROPOSALS:
        return False
    Grid, History, Ensemble = ref.Grid, ref.History, ref.Ensemble
    return True


USING_REFERENCE_CLASSES = _bind_reference_classes()
