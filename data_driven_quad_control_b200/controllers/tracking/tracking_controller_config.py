"""Tracking controller config types (reference: controllers/tracking/tracking_controller_config.py)."""

from dataclasses import dataclass
from typing import TypedDict

import torch


class TrackingControllerParams(TypedDict):
    pos_pid_gains: list[list[float]]
    kR: float


class TrackingControllerConfig(TypedDict):
    tracking_controller_params: TrackingControllerParams


@dataclass
class TrackingCtrlDroneState:
    X: torch.Tensor  # position (N, 3)
    Q: torch.Tensor  # attitude quaternion (N, 4), (w, x, y, z)
