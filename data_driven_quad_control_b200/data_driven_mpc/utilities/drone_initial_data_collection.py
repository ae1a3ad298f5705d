"""Batched initial input-output data collection for data-driven MPC controllers.

Vectorized counterpart of the reference's `collect_initial_input_output_data`
(data_driven_mpc/utilities/drone_initial_data_collection.py:32-189): every drone of the env is
stabilised at its target by the tracking controller while a persistently exciting perturbation,
uniform in `u_range`, is added to its command; the applied (clipped) input and the measured
position are recorded.  The reference collects for ONE env index with a numpy generator; here all
`num_envs` drones collect at once (one data set per future controller), entirely on the device.

The persistently exciting input can be drawn either on the device (`generator=`, one launch) or, for
stream parity with the reference, from one numpy `Generator` per env (`np_randoms=`): the draws are then
exactly the reference's (`:117-127`: one `uniform(u_range[i,0], u_range[i,1], (N, 1))` call per input,
followed by the `(N, p)` measurement-noise draw that the reference performs even when `eps_max` is 0),
so a controller built from env `b` with `np.random.default_rng(seed_b)` sees the excitation sequence the
reference's single-env collection with the same generator would have produced.
"""

from __future__ import annotations

import numpy as np
import torch

from data_driven_quad_control_b200.controllers.tracking.tracking_controller import DroneTrackingController
from data_driven_quad_control_b200.controllers.tracking.tracking_controller_config import TrackingCtrlDroneState
from data_driven_quad_control_b200.envs.hover_env import HoverEnv
from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType
from data_driven_quad_control_b200.utilities.math_utils import linear_interpolate, yaw_to_quaternion


def persistently_exciting_inputs(np_randoms, u_range, N: int, m: int, p: int, eps_max: float = 0.0):
    """(pe, w): pe (B, N, m) and w (B, N, p) float64, drawn per env with the reference's numpy calls in the
    reference's order (drone_initial_data_collection.py:117-127)."""
    u_range = np.asarray(u_range, dtype=np.float64)
    pe, w = [], []
    for rng in np_randoms:
        pe.append(np.hstack([rng.uniform(u_range[i, 0], u_range[i, 1], (N, 1)) for i in range(m)]))
        w.append(eps_max * rng.uniform(-1.0, 1.0, (N, p)))
    return np.stack(pe), np.stack(w)


def collect_initial_input_output_data_batched(
    env: HoverEnv,
    stabilizing_controller: DroneTrackingController,
    target_pos: torch.Tensor,
    target_yaw: torch.Tensor,
    input_bounds: torch.Tensor,
    u_range: torch.Tensor,
    N: int,
    generator: torch.Generator | None = None,
    np_randoms=None,
    excite_mask: torch.Tensor | None = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """Returns (u_N, y_N) with shapes (num_envs, N, m) and (num_envs, N, 3), float64.

    `input_bounds` (m, 2) are the MPC input bounds U the applied input is clipped to, `u_range`
    (m, 2) the range of the exciting perturbation (reference :117-160)."""
    if env.action_type != EnvActionType.CTBR_FIXED_YAW and env.action_type != EnvActionType.CTBR:
        raise ValueError("data collection needs a CTBR action type")
    dev, B = env.device, env.num_envs
    m = env.num_actions
    bounds = env.action_bounds
    lo = input_bounds[:, 0].to(dev, torch.float64)
    hi = input_bounds[:, 1].to(dev, torch.float64)
    ur = torch.as_tensor(u_range).to(dev, torch.float64)
    target = TrackingCtrlDroneState(X=target_pos, Q=yaw_to_quaternion(target_yaw))
    current = TrackingCtrlDroneState(X=env.get_pos(), Q=env.get_quat())
    if np_randoms is not None:
        if len(np_randoms) != B:
            raise ValueError("one numpy Generator per env is required")
        pe_np, _ = persistently_exciting_inputs(np_randoms, ur.cpu().numpy(), N, m, 3)
        pe = torch.from_numpy(pe_np).to(dev).permute(1, 0, 2).contiguous()
    else:
        pe = ur[:, 0] + (ur[:, 1] - ur[:, 0]) * torch.rand((N, B, m), device=dev, generator=generator, dtype=torch.float64)
    mask = None
    if excite_mask is not None:  # drones outside the mask are only stabilised (controller comparison: the
        pe = pe * excite_mask.to(dev, torch.float64).reshape(1, B, 1)  # other controllers' drones keep hovering)
        mask = excite_mask.to(dev, torch.bool).reshape(B, 1)
    u_N = torch.empty((B, N, m), dtype=torch.float64, device=dev)
    y_N = torch.empty((B, N, 3), dtype=torch.float64, device=dev)
    with torch.no_grad():
        for k in range(N):
            current.X, current.Q = env.get_pos(), env.get_quat()
            cmd = stabilizing_controller.compute(state_setpoint=target, state_measurement=current)[:, :m]
            # the reference accumulates in float64 (numpy) and casts to the env's float32 only for the action
            u_k = torch.minimum(torch.maximum(cmd.to(torch.float64) + pe[k], lo), hi)
            if mask is not None:  # the reference clips only the DD-MPC base env to U (:150-160); the other drones fly
                u_k = torch.where(mask, u_k, cmd.to(torch.float64))  # the raw tracking command (env bounds only)
            action = linear_interpolate(x=u_k.to(torch.float32), x_min=bounds[:, 0], x_max=bounds[:, 1], y_min=-1, y_max=1)
            env.step(torch.clamp(action, -1, 1))
            u_N[:, k] = u_k
            y_N[:, k] = env.get_pos()
    return u_N, y_N
