"""Batched data-driven MPC parameter grid search (BASELINE configs[4] on one GPU, tiny grid): every run in
flight at once, per-controller regularisation weights, reference bookkeeping and report layout."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make_env_factory(noise=1e-4):
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
    from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type

    def make_env(num_envs):
        env_cfg, obs_cfg, _, cmd_cfg = get_cfgs()
        obs_cfg["obs_noise_std"] = noise  # 1e-4 in dd_mpc_param_grid_search.py:258-274 of the reference
        env_cfg.update(episode_length_s=100000, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0,
                       termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0)  # fmt: skip
        at = EnvActionType.CTBR_FIXED_YAW
        return HoverEnv(num_envs, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device="cuda", action_type=at,
                        auto_target_updates=False, seed=11, materialize_buffers=True)  # fmt: skip

    return make_env


def test_tiny_grid_search_matches_single_run_evaluations_and_writes_the_reference_report(build_lib, tmp_path):
    from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search import batched_grid_search as gs
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.param_grid_search_config import (
        CtrlEvalStatus, DDMPCEvaluationParams, DDMPCFixedParams, DDMPCInitialDataCollectionParams, DDMPCParameterGrid,
        combinations_from_grid,
    )  # fmt: skip
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.results_writer import write_results_to_file
    from data_driven_quad_control_b200.utilities.drone_tracking_controller import create_drone_tracking_controller, hover_at_target

    make_env = _make_env_factory()
    U = np.array([[0.0, 0.52974], [-1.5708, 1.5708], [-1.5708, 1.5708]])
    fixed = DDMPCFixedParams(m=3, p=3, Q_weights=[10.0, 10.0, 10.0], R_weights=[10.0, 1.0, 1.0], S_weights=[1000.0] * 3, U=U,
                             Us=np.array([[0.001, 0.528], [-1.571, 1.570], [-1.571, 1.570]]),
                             alpha_reg_type=AlphaRegType.APPROXIMATED, ext_out_incr_in=False, n_n_mpc_step=False)  # fmt: skip
    eval_params = DDMPCEvaluationParams(eval_time_steps=12, eval_setpoints=[np.array([[0.3], [0.2], [1.8]]), np.array([[-0.2], [0.3], [1.2]])],
                                        max_target_dist_increment=0.5, num_collections_per_N=2)  # fmt: skip
    grid = DDMPCParameterGrid(N=[80], n=[2], L=[5, 6], lamb_alpha=[10.0, 50.0], lamb_sigma=[1e7], lamb_alpha_s=[1.0], lamb_sigma_s=[1e5, 1e7])
    init = DDMPCInitialDataCollectionParams(init_hover_pos=np.array([0.0, 0.0, 1.5]), u_range=np.array([[-0.1, 0.1], [-0.5, 0.5], [-0.5, 0.5]]))

    # initial data: one drone per data entry, stabilised at the hover position
    env = make_env(eval_params.num_collections_per_N)
    env.reset()
    hover = torch.tensor(init.init_hover_pos, device="cuda", dtype=torch.float32).repeat(env.num_envs, 1)
    yaw = torch.zeros(env.num_envs, device="cuda")
    env.update_target_pos(torch.arange(env.num_envs, device="cuda"), hover)
    trk = create_drone_tracking_controller(env)
    hover_at_target(env, trk, hover, yaw, min_at_target_steps=10, error_threshold=5e-2, max_steps=300)
    cache = gs.collect_data_driven_cache(env, trk, hover, yaw, U, init.u_range, grid.N, eval_params.num_collections_per_N,
                                         np.random.default_rng(0))  # fmt: skip
    assert cache.N_to_entry_indices == {80: [0, 1]} and cache.u_N[1].shape == (80, 3) and cache.drone_state[0].base_pos.shape == (1, 3)

    results = gs.batched_dd_mpc_grid_search(make_env, cache, fixed, eval_params, grid)
    combos = combinations_from_grid(grid)
    assert len(combos) == 8 and sum(len(v) for v in results.values()) == 8
    ok = results[CtrlEvalStatus.SUCCESS]
    assert len(ok) >= 6 and all(np.isfinite(r["average_RMSE"]) and 0 < r["average_RMSE"] < 1.0 for r in ok)
    assert {(r["L"], r["lamb_alpha"], r["lamb_sigma_s"]) for r in ok + results[CtrlEvalStatus.FAILURE]} == {
        (c.L, c.lamb_alpha, c.lamb_sigma_s) for c in combos}  # fmt: skip

    # one combination evaluated on its own (B = entries x setpoints = 4) gives the same per-run RMSEs as inside the
    # bucket of 16 runs: the per-controller lambda array and the run <-> env mapping are right
    c0 = combos[1]
    # (without measurement noise, whose draws depend on the env size, the runs are deterministic and independent)
    quiet = _make_env_factory(noise=0.0)
    rmse_b, failed_b = gs.evaluate_bucket(quiet, cache, [combos[0], c0], fixed, eval_params, "cuda")
    rmse_s, failed_s = gs.evaluate_bucket(quiet, cache, [c0], fixed, eval_params, "cuda")
    assert not failed_s.any() and np.array_equal(failed_b[1], failed_s[0])
    np.testing.assert_allclose(rmse_b[1], rmse_s[0], rtol=0, atol=1e-12)

    path = write_results_to_file(str(tmp_path), 12.5, 16, init, fixed, eval_params, grid, results, file_name="report.txt")
    text = open(path).read()
    for needle in ("Grid search duration: 0h 0m 12.50s", "Number of parallel processes: 16", "  alpha_reg_type: APPROXIMATED",
                   "  Total parameter combinations: 8", "  Total evaluation runs: 32", "  Evaluation runs per combination: 4",
                   f"Successful Results ({len(ok)}/8):", "    - [0.0, 0.52974]"):  # fmt: skip
        assert needle in text, needle
