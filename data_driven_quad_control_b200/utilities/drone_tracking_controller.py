"""Tracking-controller helpers (reference: utilities/drone_tracking_controller.py:32-171):
controller factory and the `hover_at_target` stabilisation loop (configs[0] workload)."""

import os

import torch

from data_driven_quad_control_b200.controllers.ctbr.ctbr_controller import DroneCTBRController
from data_driven_quad_control_b200.controllers.tracking.tracking_controller import DroneTrackingController
from data_driven_quad_control_b200.controllers.tracking.tracking_controller_config import TrackingCtrlDroneState
from data_driven_quad_control_b200.envs.hover_env import HoverEnv
from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, EnvDroneParams
from data_driven_quad_control_b200.utilities.config_utils import load_yaml_config
from data_driven_quad_control_b200.utilities.math_utils import linear_interpolate, yaw_to_quaternion

TRACKING_CONTROLLER_CONFIG_PATH = os.path.join(
    os.path.dirname(__file__), "..", "configs", "controllers", "tracking", "tracking_controller_params.yaml"
)


def create_drone_tracking_controller(env: HoverEnv) -> DroneTrackingController:
    return DroneTrackingController(
        drone_mass=EnvDroneParams.MASS,
        controller_config=load_yaml_config(TRACKING_CONTROLLER_CONFIG_PATH),
        dt=env.step_dt,
        num_envs=env.num_envs,
        device=env.device,
    )


def tracking_env_action(
    env: HoverEnv,
    tracking_controller: DroneTrackingController,
    target_state: TrackingCtrlDroneState,
    current_state: TrackingCtrlDroneState,
    ctbr_controller: DroneCTBRController | None = None,
) -> torch.Tensor:
    """One control step: tracking controller -> (external CTBR controller) -> normalised env action."""
    cmd = tracking_controller.compute(state_setpoint=target_state, state_measurement=current_state)
    if env.action_type == EnvActionType.CTBR_FIXED_YAW:
        cmd = cmd[:, :-1]  # the env keeps the yaw rate at 0
    if env.action_type == EnvActionType.ROTOR_RPMS:
        assert ctbr_controller is not None
        cmd = ctbr_controller.compute(
            rate_measurements=env.base_ang_vel, rate_setpoints=cmd[:, 1:], thrust_setpoints=cmd[:, 0]
        )
    bounds = env.action_bounds
    return linear_interpolate(x=cmd, x_min=bounds[:, 0], x_max=bounds[:, 1], y_min=-1, y_max=1)


def hover_at_target(
    env: HoverEnv,
    tracking_controller: DroneTrackingController,
    target_pos: torch.Tensor,
    target_yaw: torch.Tensor,
    min_at_target_steps: int | None = 10,
    error_threshold: float = 5e-2,
    ctbr_controller: DroneCTBRController | None = None,
    max_steps: int | None = None,
) -> int:
    """Fly every drone to `target_pos` / `target_yaw` and return once the largest position error has
    stayed below `error_threshold` for `min_at_target_steps` consecutive steps (None: forever).
    Returns the number of env steps taken.  `max_steps` (not in the reference) bounds the loop."""
    if env.action_type == EnvActionType.ROTOR_RPMS and ctbr_controller is None:
        raise ValueError("A CTBR controller must be provided when `ROTOR_RPMS` is used as the env action type.")
    target = TrackingCtrlDroneState(X=target_pos, Q=yaw_to_quaternion(target_yaw))
    current = TrackingCtrlDroneState(X=env.get_pos(), Q=env.get_quat())
    streak = steps = 0
    with torch.no_grad():
        while (min_at_target_steps is None or streak < min_at_target_steps) and (max_steps is None or steps < max_steps):
            current.X, current.Q = env.get_pos(), env.get_quat()
            env.step(tracking_env_action(env, tracking_controller, target, current, ctbr_controller))
            steps += 1
            # As in the reference (:165-171) the error is taken from `current.X` AFTER the step:
            # without observation noise `get_pos()` hands out the live position buffer, so this
            # is the post-step position; with noise it is the pre-step noisy measurement.
            pos_error = torch.abs(target.X - current.X).max().item()
            streak = streak + 1 if pos_error < error_threshold else 0
    return steps
