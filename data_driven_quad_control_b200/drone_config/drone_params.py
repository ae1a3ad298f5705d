"""Typed views of the drone parameter YAML (reference: drone_config/drone_params.py)."""

from typing import TypedDict

DroneInertia = TypedDict("DroneInertia", {k: float for k in ("Jxx", "Jxy", "Jxz", "Jyy", "Jyz", "Jzz")})


class DronePhysicalParams(TypedDict):
    mass: float
    inertia: DroneInertia


class DroneRotorParams(TypedDict):
    kf: float
    km: float
    arm_length: float
    rotor_angles_deg: list[float]
    rotor_spin_directions: list[int]


class DroneParams(TypedDict):
    drone_physical_params: DronePhysicalParams
    drone_rotor_params: DroneRotorParams
