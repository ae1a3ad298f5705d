"""Developer probe: closed-loop distance-to-target trace of the default controller on the GPU env."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from data_driven_quad_control_b200.data_driven_mpc import dd_mpc_param_grid_search as drv
from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search import batched_grid_search as gs
from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.param_grid_search_config import DDMPCCombinationParams
from data_driven_quad_control_b200.utilities.drone_tracking_controller import create_drone_tracking_controller, hover_at_target
init, fixed, evalp, grid = drv.load_dd_mpc_grid_search_params(drv.DEFAULT_CONFIG)
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType
if len(sys.argv) > 3: fixed = fixed._replace(alpha_reg_type=AlphaRegType(int(sys.argv[3])))
LAM = [float(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [50.0, 1e9, 1.0, 1e7]
noise = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-4
def make_env(n):
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
    from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type
    env_cfg, obs_cfg, _, cmd_cfg = get_cfgs(); obs_cfg["obs_noise_std"] = noise
    env_cfg.update(episode_length_s=100000, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0, termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0)
    at = EnvActionType.CTBR_FIXED_YAW
    return HoverEnv(n, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device="cuda", action_type=at, auto_target_updates=False, materialize_buffers=True)
env = make_env(3); env.reset()
hover = torch.tensor(init.init_hover_pos, device="cuda", dtype=torch.float32).repeat(3, 1); yaw = torch.zeros(3, device="cuda")
env.update_target_pos(torch.arange(3, device="cuda"), hover)
trk = create_drone_tracking_controller(env)
hover_at_target(env, trk, hover, yaw, min_at_target_steps=10, error_threshold=5e-2, max_steps=400)
cache = gs.collect_data_driven_cache(env, trk, hover, yaw, fixed.U, init.u_range, [350], 3, np.random.default_rng(0))
comb = DDMPCCombinationParams(350, 6, 35, *LAM)
import dataclasses
evalp2 = evalp._replace(eval_time_steps=int(sys.argv[2]) if len(sys.argv) > 2 else 120, max_target_dist_increment=1e9)
# trace by hand: replicate evaluate_bucket but record y
from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search import controller_evaluation as ce
orig = ce.sim_nonlinear_dd_mpc_control_loop_batched
trace = {}
def patched(env, ctrl, T, d0, inc, nn, record=False):
    r = orig(env, ctrl, T, d0, inc, nn, record=True); trace["y"] = r.y_sys.cpu().numpy(); trace["u"] = r.u_sys.cpu().numpy(); trace["yr"] = ctrl.y_r.cpu().numpy(); trace["st"] = r.solver_failed.cpu().numpy(); return r
gs.sim_nonlinear_dd_mpc_control_loop_batched = patched
rmse, failed = gs.evaluate_bucket(make_env, cache, [comb], fixed, evalp2, "cuda")
y, yr = trace["y"], trace["yr"]
dist = np.linalg.norm(y - yr[None], axis=2)
np.set_printoptions(precision=3, suppress=True, linewidth=200)
print("solver_failed", trace["st"]); print("rmse", rmse)
for t in list(range(0, dist.shape[0], 8)) + [dist.shape[0] - 1]: print(t, dist[t], "u0", trace["u"][t, 0])
