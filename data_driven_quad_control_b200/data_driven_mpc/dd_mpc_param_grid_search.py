"""Nonlinear data-driven MPC parameter grid search on the GPU (BASELINE configs[4]).

Counterpart of the reference's `data_driven_mpc/dd_mpc_param_grid_search.py` (argument parsing :120-205,
env configuration :258-274, data collection :300-380, grid search + report :385-450) without worker
processes: every evaluation run of a shape bucket is in flight at once.

  python -m data_driven_quad_control_b200.data_driven_mpc.dd_mpc_param_grid_search [--config yaml] [--seed 0]
  torchrun --nproc-per-node N --master-addr 127.0.0.1 -m ... (combinations are dealt round-robin to the ranks)
"""

from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch
import yaml

from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType
from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search import batched_grid_search as gs
from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.param_grid_search_config import (
    CtrlEvalStatus,
    DDMPCEvaluationParams,
    DDMPCFixedParams,
    DDMPCInitialDataCollectionParams,
    DDMPCParameterGrid,
)
from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.results_writer import write_results_to_file
from data_driven_quad_control_b200.envs.hover_env import HoverEnv
from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type
from data_driven_quad_control_b200.utilities.drone_tracking_controller import create_drone_tracking_controller, hover_at_target

DEFAULT_CONFIG = os.path.join(os.path.dirname(__file__), "..", "configs", "data_driven_mpc", "dd_mpc_grid_search_params.yaml")


def load_dd_mpc_grid_search_params(path: str, m: int = 3, p: int = 3):
    """(init_data_collection_params, fixed_params, eval_params, param_grid) from the reference's YAML layout
    (`grid_search_param_loader.py:30-110`)."""
    with open(path) as fh:
        cfg = yaml.safe_load(fh)
    ic, fx, ev, gr = cfg["initial_data_collection"], cfg["fixed_params"], cfg["evaluation_params"], cfg["parameter_grid"]
    init = DDMPCInitialDataCollectionParams(np.array(ic["init_hover_pos"], float), np.array(ic["u_range"], float))
    fixed = DDMPCFixedParams(m=m, p=p, Q_weights=[float(v) for v in fx["Q_weights"]], R_weights=[float(v) for v in fx["R_weights"]],
                             S_weights=[float(v) for v in fx["S_weights"]], U=np.array(fx["U"], float), Us=np.array(fx["Us"], float),
                             alpha_reg_type=AlphaRegType(fx["alpha_reg_type"]), ext_out_incr_in=bool(fx["ext_out_incr_in"]),
                             n_n_mpc_step=bool(fx["n_n_mpc_step"]))  # fmt: skip
    evalp = DDMPCEvaluationParams(int(ev["eval_time_steps"]), [np.array(s, float).reshape(-1, 1) for s in ev["eval_setpoints"]],
                                  float(ev["max_target_dist_increment"]), int(ev["num_collections_per_N"]))  # fmt: skip
    grid = DDMPCParameterGrid(*[[type(v)(v) for v in gr[k]] for k in DDMPCParameterGrid._fields])
    return init, fixed, evalp, grid


def make_env_factory(device):
    def make_env(num_envs: int) -> HoverEnv:
        env_cfg, obs_cfg, _, cmd_cfg = get_cfgs()
        obs_cfg["obs_noise_std"] = 1e-4  # reference dd_mpc_param_grid_search.py:258-274
        env_cfg.update(episode_length_s=100000, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0,
                       termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0)  # fmt: skip
        at = EnvActionType.CTBR_FIXED_YAW
        return HoverEnv(num_envs, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device=device, action_type=at,
                        auto_target_updates=False, materialize_buffers=True, independent_envs=True)  # fmt: skip
        # independent_envs: the reference steps ONE env with num_envs = num_processes drones (dd_mpc_param_grid_search.py:
        # 282-291), so a crashed drone zeroes the rate PID of the few runs that happen to be evaluated beside it.  With
        # thousands of runs in one batch every crash would hit every run: the runs are decoupled instead (each behaves as
        # with num_processes = 1), which also makes a run's score independent of its neighbours

    return make_env


def run_grid_search(config: str = DEFAULT_CONFIG, seed: int = 0, device="cuda", output_dir: str | None = None) -> dict:
    init, fixed, evalp, grid = load_dd_mpc_grid_search_params(config)
    make_env = make_env_factory(device)
    # the measurement noise of `HoverEnv.get_pos(add_noise=True)` comes from torch's global generator (as in the reference):
    # seeded here, two runs with the same seed evaluate the same noise realisation and write the same report
    torch.manual_seed(seed)
    t0 = time.perf_counter()
    env = make_env(evalp.num_collections_per_N)
    env.reset()
    B = env.num_envs
    hover = torch.tensor(init.init_hover_pos, device=env.device, dtype=torch.float32).repeat(B, 1)
    yaw = torch.zeros(B, device=env.device)
    env.update_target_pos(torch.arange(B, device=env.device), hover)
    trk = create_drone_tracking_controller(env)
    hover_at_target(env, trk, hover, yaw, min_at_target_steps=10, error_threshold=5e-2, max_steps=400)
    cache = gs.collect_data_driven_cache(env, trk, hover, yaw, fixed.U, init.u_range, grid.N, evalp.num_collections_per_N,
                                         np.random.default_rng(seed))  # every rank collects the same data (same seed)
    t1 = time.perf_counter()
    results = gs.batched_dd_mpc_grid_search(make_env, cache, fixed, evalp, grid, device)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n_comb = sum(len(v) for v in results.values())
    runs = n_comb * evalp.num_collections_per_N * len(evalp.eval_setpoints)
    summary = {"combinations": n_comb, "successful": len(results[CtrlEvalStatus.SUCCESS]), "evaluation_runs": runs,
               "closed_loop_steps": runs * evalp.eval_time_steps, "data_collection_s": t1 - t0, "grid_search_s": t2 - t1}  # fmt: skip
    ok = sorted(results[CtrlEvalStatus.SUCCESS], key=lambda r: r["average_RMSE"])
    summary["best"] = ok[0] if ok else None
    rank, _ = (torch.distributed.get_rank(), 0) if torch.distributed.is_initialized() else (0, 0)
    if output_dir and rank == 0:
        summary["report"] = write_results_to_file(output_dir, t2 - t0, runs, init, fixed,
                                                  evalp, grid, results)  # fmt: skip
    return summary


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--output-dir", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    summary = run_grid_search(args.config, args.seed, torch.device("cuda", local), args.output_dir)
    if int(os.environ.get("RANK", "0")) == 0:
        import json

        print(json.dumps(summary))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
