"""Developer probe: long closed loops on the GPU - no NaN / inf, no failed controller, statistics consistent.
  (a) 65 536 envs x 5000 random-action steps (auto target updates, resets): every state finite, env_steps / resets totals
  (b) 256 DD-MPC controllers x 400 closed-loop steps on the env (two setpoints), default mode and alpha_reg_type PREVIOUS"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import bench
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType, BatchedNonlinearDataDrivenMPC
from data_driven_quad_control_b200.data_driven_mpc.utilities.drone_initial_data_collection import collect_initial_input_output_data_batched
from data_driven_quad_control_b200.envs.hover_env import HoverEnv
from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type
from data_driven_quad_control_b200.utilities.drone_tracking_controller import create_drone_tracking_controller, hover_at_target
from data_driven_quad_control_b200.utilities.math_utils import linear_interpolate
dev = torch.device("cuda"); out = {}
N, T = 1 << 16, 5000
env = bench.make_env(N, dev, 3); env.reset()
acts = torch.rand((64, N, 3), device=dev) * 2.4 - 1.2
for t in range(T): env.step(acts[t % 64])
st = env.rollout_stats()
fin = all(bool(torch.isfinite(x).all()) for x in (env.base_pos, env.base_quat, env.obs_buf, env.rew_buf))
out["env"] = {"steps": T, "envs": N, "finite": fin, "env_steps": st["env_steps"], "resets": st["resets"], "quat_norm_err": float((env.base_quat.norm(dim=1) - 1).abs().max())}
P = bench.MPC_DEFAULT; B = 256; ell = P["L"] + P["n"] + 1
for name, art in (("approximated", AlphaRegType.APPROXIMATED), ("previous", AlphaRegType.PREVIOUS)):
    env_cfg, obs_cfg, _, cmd_cfg = get_cfgs(); obs_cfg["obs_noise_std"] = 1e-4
    env_cfg.update(episode_length_s=100000, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0, termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0)
    at = EnvActionType.CTBR_FIXED_YAW
    e = HoverEnv(B, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device=dev, action_type=at, auto_target_updates=False, seed=11, materialize_buffers=True, independent_envs=True)
    e.reset(); hover = torch.tensor(P["init_hover_pos"], device=dev).repeat(B, 1); yaw = torch.zeros(B, device=dev)
    e.update_target_pos(torch.arange(B, device=dev), hover); trk = create_drone_tracking_controller(e)
    hover_at_target(e, trk, hover, yaw, min_at_target_steps=10, error_threshold=5e-2, max_steps=300)
    u_N, y_N = collect_initial_input_output_data_batched(e, trk, hover, yaw, torch.tensor(P["U"]), torch.tensor(P["u_range"]), P["N"], torch.Generator(device=dev).manual_seed(5))
    c = BatchedNonlinearDataDrivenMPC(n=P["n"], m=3, p=3, u=u_N, y=y_N, L=P["L"], Q=np.kron(np.eye(ell), np.diag(P["Q"])), R=np.kron(np.eye(ell), np.diag(P["R"])), S=np.diag(P["S"]),
        y_r=np.array(P["y_r"]), lamb_alpha=P["lamb_alpha"], lamb_sigma=P["lamb_sigma"], U=P["U"], Us=P["Us"], alpha_reg_type=art, lamb_alpha_s=P["lamb_alpha_s"], lamb_sigma_s=P["lamb_sigma_s"], device=dev, solve_on_init=False)
    bad = 0; dist = []
    for sp in ([1.0, 1.0, 2.5], [0.0, 0.0, 1.5]):
        y_r = torch.tensor(sp, device=dev).repeat(B, 1); e.update_target_pos(torch.arange(B, device=dev), y_r); c.set_output_setpoint(np.tile(np.array(sp)[None], (B, 1)))
        for t in range(200):
            c.update_and_solve_data_driven_mpc(); bad += int((c.status != 0).sum())
            u0 = c.get_optimal_control_input_at_step(0)
            e.step(linear_interpolate(x=u0.to(torch.float32), x_min=e.action_bounds[:, 0], x_max=e.action_bounds[:, 1], y_min=-1, y_max=1))
            c.store_input_output_measurement(u0, e.get_pos())
        dist.append(torch.linalg.norm(e.base_pos - y_r, dim=1))
    out["ddmpc_" + name] = {"controllers": B, "steps": 400, "failed_solves": bad, "finite": bool(torch.isfinite(e.base_pos).all()),
                            "final_dist_median_m": [float(d.median()) for d in dist], "final_dist_p90_m": [float(d.quantile(0.9)) for d in dist],
                            "within_0.1m_frac": [float((d < 0.1).float().mean()) for d in dist]}
print(json.dumps(out))
