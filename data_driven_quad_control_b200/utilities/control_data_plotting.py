"""`ControlTrajectory`: the record the reference's comparison run produces and its plotting consumes
(utilities/control_data_plotting.py:15-19).  Plotting itself is out of scope here."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class ControlTrajectory:
    control_inputs: list[np.ndarray]  # per drone: (T, m) control inputs in physical units
    system_outputs: list[np.ndarray]  # per drone: (T, 3) true positions
    system_setpoint: np.ndarray  # (T, 3) target position at every step
