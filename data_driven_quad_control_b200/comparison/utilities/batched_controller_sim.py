"""Tracking controller vs trained policy vs data-driven MPC on the same targets, all drones in one env.

Device-resident counterpart of the reference's `comparison/utilities/parallel_controller_sim.py:60-443`
(three controller worker processes exchanging actions / observations with the env process through queues):
env rows [0, R) fly the SE(3) tracking controller, [R, 2R) the policy, [2R, 3R) the data-driven MPC
controllers (R replicas each; the reference runs R = 1).  Sequence as in the reference (:150-340):
stabilise every drone at `init_hover_pos`, collect the DD-MPC initial data on its drones while the others keep
hovering, then visit `eval_setpoints`, `steps_per_setpoint` steps each, recording the clamped actions converted
to physical units and the true positions (`construct_trajectory_data`, :399-443).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from data_driven_quad_control_b200 import _lib
from data_driven_quad_control_b200.controllers.tracking.tracking_controller_config import TrackingCtrlDroneState
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import (
    BatchedNonlinearDataDrivenMPC,
)
from data_driven_quad_control_b200.data_driven_mpc.utilities.drone_initial_data_collection import (
    collect_initial_input_output_data_batched,
)
from data_driven_quad_control_b200.envs.hover_env import HoverEnv
from data_driven_quad_control_b200.utilities.control_data_plotting import ControlTrajectory
from data_driven_quad_control_b200.utilities.drone_tracking_controller import create_drone_tracking_controller, hover_at_target
from data_driven_quad_control_b200.utilities.math_utils import linear_interpolate, yaw_to_quaternion


def batched_controller_comparison(env: HoverEnv, policy, dd_mpc_kwargs: dict, u_range, init_hover_pos, eval_setpoints,
                                  steps_per_setpoint: int, np_random: np.random.Generator) -> ControlTrajectory:  # fmt: skip
    """`env.num_envs` must be 3 R.  `policy` maps (R, num_obs) observations to (R, num_actions) actions
    (`FusedActorMLP`); `dd_mpc_kwargs` are the controller keywords except `u`, `y`, `y_r` (n, m, p, L, Q, R, S,
    lamb_*, U, Us, alpha_reg_type, ...)."""
    if env.num_envs % 3:
        raise ValueError("the comparison env needs 3 R drones (tracking | policy | data-driven MPC)")
    R, dev, m = env.num_envs // 3, env.device, env.num_actions
    lib = _lib.load()
    idx = torch.arange(env.num_envs, device=dev)
    bounds = env.action_bounds
    env.reset()
    hover = torch.tensor(np.asarray(init_hover_pos, np.float32), device=dev).repeat(env.num_envs, 1)
    yaw = torch.zeros(env.num_envs, device=dev)
    env.update_target_pos(idx, hover)
    trk = create_drone_tracking_controller(env)
    hover_at_target(env, trk, hover, yaw, min_at_target_steps=10, error_threshold=5e-2, max_steps=400)
    mask = torch.zeros(env.num_envs, device=dev)
    mask[2 * R :] = 1.0
    N = int(dd_mpc_kwargs["N"]) if "N" in dd_mpc_kwargs else 350
    kw = {k: v for k, v in dd_mpc_kwargs.items() if k != "N"}
    u_N, y_N = collect_initial_input_output_data_batched(
        env, trk, hover, yaw, torch.as_tensor(np.asarray(kw["U"])), torch.as_tensor(np.asarray(u_range)), N,
        np_randoms=np_random.spawn(env.num_envs), excite_mask=mask,
    )  # fmt: skip
    y_r0 = np.tile(np.asarray(eval_setpoints[0], np.float64).reshape(1, -1), (R, 1))
    mpc = BatchedNonlinearDataDrivenMPC(u=u_N[2 * R :], y=y_N[2 * R :], y_r=y_r0, device=dev, solve_on_init=False, **kw)
    cfg = C.byref(mpc._cfg)
    u_applied = torch.empty((R, m), dtype=torch.float64, device=dev)
    actions = torch.zeros((env.num_envs, m), dtype=torch.float32, device=dev)
    target_state = TrackingCtrlDroneState(X=hover.clone(), Q=yaw_to_quaternion(yaw))
    current = TrackingCtrlDroneState(X=env.get_pos(), Q=env.get_quat())
    act_log, pos_log, tgt_log = [], [], []
    obs, _ = env.get_observations()
    with torch.no_grad():
        for sp in eval_setpoints:
            tgt = torch.tensor(np.asarray(sp, np.float32).reshape(1, 3), device=dev).repeat(env.num_envs, 1)
            env.update_target_pos(idx, tgt)
            target_state.X = tgt
            mpc.set_output_setpoint(np.tile(np.asarray(sp, np.float64).reshape(1, -1), (R, 1)))
            for _ in range(steps_per_setpoint):
                # tracking controller rows (the yaw-rate command is dropped for CTBR_FIXED_YAW, as in the reference)
                current.X, current.Q = env.get_pos(), env.get_quat()
                cmd = trk.compute(state_setpoint=target_state, state_measurement=current)[:, :m]
                actions[:R] = linear_interpolate(x=cmd[:R], x_min=bounds[:, 0], x_max=bounds[:, 1], y_min=-1, y_max=1)
                # policy rows
                actions[R : 2 * R] = policy.act_inference(obs[R : 2 * R].contiguous())
                # data-driven MPC rows: solve, then optimal input -> normalised action straight into the action buffer
                mpc.update_and_solve_data_driven_mpc()
                with _lib.device_guard(dev):
                    rc = lib.ddqc_ddmpc_action(cfg, mpc.optimal_u.data_ptr(), 0, bounds.contiguous().data_ptr(),
                                               actions[2 * R :].data_ptr(), u_applied.data_ptr(), R, _lib.stream_ptr(dev))  # fmt: skip
                _lib.check(rc, "ddqc_ddmpc_action")
                act_log.append(actions.clone())
                obs, *_ = env.step(actions)
                mpc.store_input_output_measurement(u_applied, env.get_pos()[2 * R :])
                pos_log.append(env.get_pos(add_noise=False).clone())
                tgt_log.append(np.asarray(sp, np.float64).reshape(-1))
    env.check_ground_plane("controller comparison")  # warns if a drone sank below the (unmodelled) floor
    # construct_trajectory_data of the reference: clamp to [-1, 1] (the env clips), back to physical units
    a = torch.stack(act_log).clamp(-1, 1)
    u_phys = linear_interpolate(x=a, x_min=-1, x_max=1, y_min=bounds[:, 0], y_max=bounds[:, 1]).cpu().numpy().transpose(1, 0, 2)
    pos = torch.stack(pos_log).cpu().numpy().transpose(1, 0, 2)
    return ControlTrajectory(control_inputs=[u_phys[i] for i in range(u_phys.shape[0])],
                             system_outputs=[pos[i] for i in range(pos.shape[0])], system_setpoint=np.vstack(tgt_log))  # fmt: skip
