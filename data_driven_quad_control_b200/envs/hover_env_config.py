"""Environment configuration surface (reference: envs/hover_env_config.py).

Same names, values and semantics as the reference module: `get_cfgs()` default dicts
(:34-94), `EnvDroneParams` / `EnvCTBRControllerConfig` YAML holders (:98-118, :184-194),
`EnvActionType` (:122-168), `EnvActionBounds` (:172-180) and the `EnvState` snapshot (:198-226).
"""

from __future__ import annotations

import os
from dataclasses import dataclass, fields
from enum import Enum
from typing import Any, Optional

import torch

from data_driven_quad_control_b200.controllers.ctbr.ctbr_controller_config import CTBRControllerConfig
from data_driven_quad_control_b200.drone_config.drone_params import DroneParams
from data_driven_quad_control_b200.utilities.config_utils import load_yaml_config
from data_driven_quad_control_b200.utilities.vectorized_pid_controller import VectorizedControllerState

_CONFIG_DIR = os.path.join(os.path.dirname(__file__), "..", "configs")
DRONE_PARAMS_PATH = os.path.join(_CONFIG_DIR, "drone", "cf2x_drone_params.yaml")
CTBR_CONTROLLER_CONFIG_PATH = os.path.join(_CONFIG_DIR, "controllers", "ctbr", "ctbr_controller_params.yaml")

CfgDict = dict[str, Any]


def get_cfgs() -> tuple[CfgDict, CfgDict, CfgDict, CfgDict]:
    """Default (env_cfg, obs_cfg, reward_cfg, command_cfg)."""
    env_cfg: CfgDict = dict(
        dt=0.01,  # 100 Hz physics
        decimation=4,  # 25 Hz control
        simulate_action_latency=True,
        clip_actions=1.0,
        actuator_noise_std=0.0,  # RPM
        termination_if_roll_greater_than=180,  # deg
        termination_if_pitch_greater_than=180,
        termination_if_close_to_ground=0.1,
        termination_if_x_greater_than=3.0,
        termination_if_y_greater_than=3.0,
        termination_if_z_greater_than=2.0,
        termination_if_ang_vel_greater_than=20,  # rad/s
        termination_if_lin_vel_greater_than=20,  # m/s
        base_init_pos=[0.0, 0.0, 1.0],
        base_init_quat=[1.0, 0.0, 0.0, 0.0],
        episode_length_s=15.0,
        at_target_threshold=0.1,
        min_hover_time_s=0.5,
        resampling_time_s=3.0,
        visualize_target=False,
        visualize_camera=False,
        max_visualize_FPS=100,
    )
    obs_cfg: CfgDict = dict(
        obs_scales=dict(rel_pos=1 / 3.0, lin_vel=1 / 3.0, ang_vel=1 / 3.14159),
        obs_noise_std=0.0,
    )
    reward_cfg: CfgDict = dict(
        yaw_lambda=-10.0,
        reward_scales=dict(
            target=10.0, closeness=1.5, hover_time=0.01, smooth=-1e-4, yaw=0.01, angular=-2e-4, crash=-10.0
        ),
    )
    command_cfg: CfgDict = dict(
        num_commands=3, pos_x_range=[-1.0, 1.0], pos_y_range=[-1.0, 1.0], pos_z_range=[1.0, 3.0]
    )
    return env_cfg, obs_cfg, reward_cfg, command_cfg


class EnvDroneParams:
    """Drone physical / rotor parameters loaded once from YAML."""

    _params: DroneParams = load_yaml_config(DRONE_PARAMS_PATH)
    _phys = _params["drone_physical_params"]
    _rotor = _params["drone_rotor_params"]

    MASS = _phys["mass"]
    WEIGHT = MASS * 9.81  # N
    INERTIA = _phys["inertia"]
    KF = _rotor["kf"]
    KM = _rotor["km"]
    ARM_LENGTH = _rotor["arm_length"]
    ROTOR_ANGLES_DEG = _rotor["rotor_angles_deg"]
    ROTOR_SPIN_DIRECTIONS = _rotor["rotor_spin_directions"]

    @classmethod
    def get(cls) -> DroneParams:
        return cls._params


class EnvActionType(Enum):
    ROTOR_RPMS = 0  # [RPM_1..RPM_4]
    CTBR = 1  # [F_z, w_x, w_y, w_z]
    CTBR_FIXED_YAW = 2  # [F_z, w_x, w_y], w_z setpoint fixed at 0

    @classmethod
    def get_num_actions(cls, action_type: "EnvActionType") -> int:
        table = {cls.ROTOR_RPMS: 4, cls.CTBR: 4, cls.CTBR_FIXED_YAW: 3}
        if action_type not in table:
            raise ValueError(f"Unknown action_type: {action_type}")
        return table[action_type]

    @classmethod
    def get_num_obs(cls, action_type: "EnvActionType") -> int:
        table = {cls.ROTOR_RPMS: 17, cls.CTBR: 17, cls.CTBR_FIXED_YAW: 16}
        if action_type not in table:
            raise ValueError(f"Unknown action_type: {action_type}")
        return table[action_type]


class EnvActionBounds:
    BASE_RPM = 14468.429183500699  # hover RPM
    MIN_RPM = 0.2 * BASE_RPM
    MAX_RPM = 1.8 * BASE_RPM
    MAX_THRUST = 2 * EnvDroneParams.WEIGHT  # N
    MAX_ANG_VELS = [15.0, 15.0, 10.0]  # roll, pitch, yaw rate limits, rad/s


class EnvCTBRControllerConfig:
    _config: CTBRControllerConfig = load_yaml_config(CTBR_CONTROLLER_CONFIG_PATH)
    RATE_PID_GAINS = _config["ctbr_controller_params"]["rate_pid_gains"]

    @classmethod
    def get(cls) -> CTBRControllerConfig:
        return cls._config


@dataclass
class EnvState:
    """Snapshot of the env state for a set of env indices (+ the whole-batch CTBR PID state)."""

    base_pos: torch.Tensor
    base_quat: torch.Tensor
    base_lin_vel: torch.Tensor
    base_ang_vel: torch.Tensor
    commands: torch.Tensor
    episode_length: torch.Tensor
    last_actions: torch.Tensor
    ctbr_controller_state: Optional[VectorizedControllerState] = None

    def to(self, device: torch.device | str) -> "EnvState":
        moved = {}
        for f in fields(self):
            val = getattr(self, f.name)
            if f.name == "ctbr_controller_state":
                moved[f.name] = val.to(device) if val else None
            else:
                moved[f.name] = val.to(device)
        return EnvState(**moved)

    def to_cpu(self) -> "EnvState":
        return self.to("cpu")
