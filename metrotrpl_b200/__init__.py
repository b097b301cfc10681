"""metrotrpl_b200 - B200-native forward simulation + likelihood path for MetroTRPL.

Only the data-parallel hot path lives here: batched drift-diffusion forward
solves, PL/TRTS readout and per-curve log-likelihoods on the GPU behind the reference's own
solve / eval_trial_move / metro / dense-sampling call shapes.
"""
from . import _capi  # noqa: F401

__all__ = ["_capi"]
__version__ = "0.1.0"
