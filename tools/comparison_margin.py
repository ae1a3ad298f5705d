"""Developer probe: distance to the setpoints at the end of each comparison leg for several torch seeds
(`HoverEnv.get_pos(add_noise=True)` draws from torch's global generator, as the reference does)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from data_driven_quad_control_b200.comparison.utilities.batched_controller_sim import batched_controller_comparison
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType
from data_driven_quad_control_b200.envs.hover_env import HoverEnv
from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type
from data_driven_quad_control_b200.learning.policy import FusedActorMLP

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 12
g = np.load(os.path.join(ROOT, "tests", "golden", "policy_demo.npz"))
w = {f"actor.{k}.{n}": torch.from_numpy(g[f"W{k}" + ("" if n == "weight" else "_b")]) for k in (0, 2, 4) for n in ("weight", "bias")}
policy = FusedActorMLP(w)
P = bench.MPC_DEFAULT
ell = P["L"] + P["n"] + 1
kw = dict(n=P["n"], m=3, p=3, L=P["L"], N=P["N"], Q=np.kron(np.eye(ell), np.diag(P["Q"])), R=np.kron(np.eye(ell), np.diag(P["R"])),
          S=np.diag(P["S"]), lamb_alpha=P["lamb_alpha"], lamb_sigma=P["lamb_sigma"], U=P["U"], Us=P["Us"],
          alpha_reg_type=AlphaRegType.APPROXIMATED, lamb_alpha_s=P["lamb_alpha_s"], lamb_sigma_s=P["lamb_sigma_s"])
setpoints, steps = [[1.0, -1.0, 2.5], [0.0, 0.0, 1.5]], 175
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import BatchedNonlinearDataDrivenMPC as _M
_orig = _M.update_and_solve_data_driven_mpc
_log = []
def _patched(self):
    r = _orig(self)
    _log.append((self.status.cpu().numpy().copy(), self.pivoting_passes.cpu().numpy().copy(), self.ipm_iterations.cpu().numpy().copy()))
    return r
_M.update_and_solve_data_driven_mpc = _patched
only = int(os.environ["ONLY_SEED"]) if "ONLY_SEED" in os.environ else None
for seed in ([only] if only is not None else range(seeds)):
    _log.clear()
    torch.manual_seed(seed)
    env_cfg, obs_cfg, _, cmd_cfg = get_cfgs()
    obs_cfg["obs_noise_std"] = 1e-4
    env_cfg.update(episode_length_s=100000, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0,
                   termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0)
    at = EnvActionType.CTBR_FIXED_YAW
    env = HoverEnv(3 * R, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device="cuda", action_type=at,
                   auto_target_updates=False, seed=5, materialize_buffers=True)
    traj = batched_controller_comparison(env, policy, kw, P["u_range"], P["init_hover_pos"], setpoints, steps, np.random.default_rng(0))
    d = [(round(float(np.linalg.norm(traj.system_outputs[i][steps - 1] - np.array(setpoints[0]))), 3),
          round(float(np.linalg.norm(traj.system_outputs[i][-1] - np.array(setpoints[1]))), 3)) for i in range(3 * R)]
    st = np.array([l[0] for l in _log]); pp = np.array([l[1] for l in _log]); ii = np.array([l[2] for l in _log])
    print(seed, d, "status!=0:", int((st != 0).sum()), "max passes", int(pp.max()), "ipm solves", int((ii > 0).sum()), flush=True)
    if only is not None:
        for i in range(2 * R, 3 * R):
            dist = np.linalg.norm(traj.system_outputs[i] - traj.system_setpoint, axis=1)
            print("ctrl", i, "dist every 25 steps:", dist[::25].round(3), "final", dist[-1].round(3))
            print("  status nonzero steps:", np.nonzero(st[:, i - 2 * R])[0][:20], "ipm steps:", np.nonzero(ii[:, i - 2 * R])[0][:20], "passes max", pp[:, i - 2 * R].max())
            print("  inputs last 5:", traj.control_inputs[i][-5:].round(4))
