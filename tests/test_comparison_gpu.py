"""Three-way controller comparison (tracking controller | trained policy | data-driven MPC) in one env:
BASELINE configs[4], second half (configs/comparison/controller_comparison_config.yaml:25-38 of the reference)."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_three_controllers_reach_the_setpoints(build_lib):
    import bench
    from data_driven_quad_control_b200.comparison.utilities.batched_controller_sim import batched_controller_comparison
    from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
    from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type
    from data_driven_quad_control_b200.learning.policy import FusedActorMLP

    env_cfg, obs_cfg, _, cmd_cfg = get_cfgs()
    obs_cfg["obs_noise_std"] = 1e-4
    env_cfg.update(episode_length_s=100000, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0,
                   termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0)  # fmt: skip
    at = EnvActionType.CTBR_FIXED_YAW
    R = 2
    env = HoverEnv(3 * R, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device="cuda", action_type=at,
                   auto_target_updates=False, seed=5, materialize_buffers=True)  # fmt: skip
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "policy_demo.npz"))  # the reference's trained actor
    w = {f"actor.{k}.{n}": torch.from_numpy(g[f"W{k}" + ("" if n == "weight" else "_b")]) for k in (0, 2, 4) for n in ("weight", "bias")}
    policy = FusedActorMLP(w)
    P = bench.MPC_DEFAULT
    ell = P["L"] + P["n"] + 1
    kw = dict(n=P["n"], m=3, p=3, L=P["L"], N=P["N"], Q=np.kron(np.eye(ell), np.diag(P["Q"])), R=np.kron(np.eye(ell), np.diag(P["R"])),
              S=np.diag(P["S"]), lamb_alpha=P["lamb_alpha"], lamb_sigma=P["lamb_sigma"], U=P["U"], Us=P["Us"],
              alpha_reg_type=AlphaRegType.APPROXIMATED, lamb_alpha_s=P["lamb_alpha_s"], lamb_sigma_s=P["lamb_sigma_s"])  # fmt: skip
    setpoints, steps = [[1.0, -1.0, 2.5], [0.0, 0.0, 1.5]], 175
    traj = batched_controller_comparison(env, policy, kw, P["u_range"], P["init_hover_pos"], setpoints, steps, np.random.default_rng(0))
    assert len(traj.control_inputs) == len(traj.system_outputs) == 3 * R
    assert traj.control_inputs[0].shape == (2 * steps, 3) and traj.system_outputs[0].shape == (2 * steps, 3)
    assert traj.system_setpoint.shape == (2 * steps, 3) and (traj.system_setpoint[0] == setpoints[0]).all()
    lo, hi = env.action_bounds[:, 0].cpu().numpy(), env.action_bounds[:, 1].cpu().numpy()
    for i in range(3 * R):
        assert (traj.control_inputs[i] >= lo - 1e-6).all() and (traj.control_inputs[i] <= hi + 1e-6).all()
        d1 = np.linalg.norm(traj.system_outputs[i][steps - 1] - np.array(setpoints[0]))
        d2 = np.linalg.norm(traj.system_outputs[i][-1] - np.array(setpoints[1]))
        # every controller (PID tracking, the reference's trained policy, DD-MPC with the shipped defaults) gets there
        assert d1 < 0.25 and d2 < 0.25, (i, d1, d2)
