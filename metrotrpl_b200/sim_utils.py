"""Grid and constants of the reference's sim_utils.py that the hot path touches.

Mirrors /root/reference/sim_utils.py:13-23 (constants) and :248-283 (Grid).
"""
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
