"""Batched closed-loop evaluator (ddqc_ddmpc_action / ddqc_ddmpc_post_step) against the numpy restatement
of the reference's sim_nonlinear_dd_mpc_control_loop_parallel (oracle/ddmpc_eval_oracle.py)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(B, n_mpc_step=1, seed=3, alpha_reg_type=None, **ctor_extra):
    import ddmpc_oracle as mo
    from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import (
        BatchedNonlinearDataDrivenMPC,
    )
    from data_driven_quad_control_b200.data_driven_mpc.utilities.drone_initial_data_collection import (
        collect_initial_input_output_data_batched,
    )
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
    from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type
    from data_driven_quad_control_b200.utilities.drone_tracking_controller import (
        create_drone_tracking_controller,
        hover_at_target,
    )

    dev = torch.device("cuda")
    prm = mo.default_params()
    prm.n, prm.L, prm.N = 2, 5, 80
    if alpha_reg_type is not None:
        prm.alpha_reg_type = alpha_reg_type
    env_cfg, obs_cfg, _, cmd_cfg = get_cfgs()
    obs_cfg["obs_noise_std"] = 1e-4
    env_cfg.update(episode_length_s=100, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0,
                   termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0)  # fmt: skip
    at = EnvActionType.CTBR_FIXED_YAW
    env = HoverEnv(B, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device=dev, action_type=at,
                   auto_target_updates=False, seed=seed, materialize_buffers=True)  # fmt: skip
    env.reset()
    hover = torch.tensor([0.0, 0.0, 1.5], device=dev).repeat(B, 1)
    yaw = torch.zeros(B, device=dev)
    env.update_target_pos(torch.arange(B, device=dev), hover)
    trk = create_drone_tracking_controller(env)
    hover_at_target(env, trk, hover, yaw, min_at_target_steps=10, error_threshold=5e-2, max_steps=300)
    u_range = np.array([[-0.1, 0.1], [-0.5, 0.5], [-0.5, 0.5]])
    u_N, y_N = collect_initial_input_output_data_batched(
        env, trk, hover, yaw, torch.tensor(prm.U), torch.tensor(u_range), prm.N,
        np_randoms=[np.random.default_rng(seed + b) for b in range(B)],  # the reference's numpy stream, one per env
    )
    # excitation = the reference's draws; applied input = clip(stabilising command (fp32) + excitation, U)
    for b in range(B):
        rng = np.random.default_rng(seed + b)
        pe = np.hstack([rng.uniform(u_range[i, 0], u_range[i, 1], (prm.N, 1)) for i in range(3)])
        u_b = u_N[b].cpu().numpy()
        assert (u_b >= prm.U[:, 0] - 1e-15).all() and (u_b <= prm.U[:, 1] + 1e-15).all()
        inside = (u_b > prm.U[:, 0]) & (u_b < prm.U[:, 1])
        cmd = (u_b - pe)[inside]
        assert np.abs(cmd - cmd.astype(np.float32)).max() < 1e-9  # what is left is the fp32 controller command
    y_r = np.tile(np.array([[0.3, 0.2, 1.8]]), (B, 1)) + 0.05 * np.arange(B)[:, None]
    kw = mo.ctor_kwargs(prm, u_N[0].cpu().numpy(), y_N[0].cpu().numpy(), y_r[0])
    kw.update(u=u_N, y=y_N, y_r=y_r, n_mpc_step=n_mpc_step, **ctor_extra)
    ctrl = BatchedNonlinearDataDrivenMPC(**kw, device=dev, solve_on_init=False)
    env.update_target_pos(torch.arange(B, device=dev), torch.from_numpy(y_r).to(dev))
    return prm, env, ctrl, u_N.cpu().numpy(), y_N.cpu().numpy(), y_r


@pytest.mark.parametrize("n_n_mpc_step", [False, True])
def test_batched_closed_loop_matches_reference_loop(build_lib, n_n_mpc_step):
    import ddmpc_eval_oracle as eo
    import ddmpc_oracle as mo
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.controller_evaluation import (
        sim_nonlinear_dd_mpc_control_loop_batched,
    )

    B, T = 3, 6
    prm, env, ctrl, u0, y0, y_r = _setup(B, n_mpc_step=2 if n_n_mpc_step else 1)
    d0 = np.linalg.norm(env.get_pos(add_noise=False).cpu().numpy().astype(np.float64) - y_r, axis=1)
    res = sim_nonlinear_dd_mpc_control_loop_batched(env, ctrl, T, torch.from_numpy(d0), None, n_n_mpc_step, record=True)
    assert not bool(res.failed.any())
    y_sys, u_sys = res.y_sys.cpu().numpy(), res.u_sys.cpu().numpy()
    for b in range(B):
        # the reference loop on the literal-QP oracle controller, fed with the SAME measurements
        replay = iter(y_sys[:, b])
        orc = mo.NonlinearDDMPCOracle(**{**mo.ctor_kwargs(prm, u0[b], y0[b], y_r[b]), "n_mpc_step": ctrl.n_mpc_step}, solve_on_init=False)
        rmse, u_ref, _ = eo.sim_control_loop(orc, lambda u: next(replay), T, d0[b], None, n_n_mpc_step)
        np.testing.assert_allclose(u_sys[:, b], u_ref, rtol=0, atol=2e-5)
        np.testing.assert_allclose(float(res.rmse[b]), rmse, rtol=1e-13, atol=0)
        # rolling windows: initial data shifted by T samples, then the applied inputs / measured outputs
        np.testing.assert_array_equal(ctrl.u[b].cpu().numpy(), np.vstack([u0[b][T:], u_sys[:, b]]))
        np.testing.assert_array_equal(ctrl.y[b].cpu().numpy(), np.vstack([y0[b][T:], y_sys[:, b]]))


@pytest.mark.parametrize("alpha_reg_type,thr", [(1, None), (0, 1e9), (1, 1e9)])  # 1e9: the data freezes after the first solve
def test_batched_closed_loop_optional_modes(build_lib, alpha_reg_type, thr):
    """The evaluator with the controller modes of SURVEY 8 a16 switched on (alpha_reg_type PREVIOUS, update_cost_threshold;
    dual recovery after every solve), replayed through the literal-QP oracle controller in the same mode."""
    import ddmpc_eval_oracle as eo
    import ddmpc_oracle as mo
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.controller_evaluation import (
        sim_nonlinear_dd_mpc_control_loop_batched,
    )

    B, T = 3, 8
    prm, env, ctrl, u0, y0, y_r = _setup(B, alpha_reg_type=alpha_reg_type, update_cost_threshold=thr)
    d0 = np.linalg.norm(env.get_pos(add_noise=False).cpu().numpy().astype(np.float64) - y_r, axis=1)
    res = sim_nonlinear_dd_mpc_control_loop_batched(env, ctrl, T, torch.from_numpy(d0), None, False, record=True)
    assert not bool(res.failed.any())
    y_sys, u_sys = res.y_sys.cpu().numpy(), res.u_sys.cpu().numpy()
    frozen = 0
    for b in range(B):
        replay = iter(y_sys[:, b])
        orc = mo.NonlinearDDMPCOracle(**{**mo.ctor_kwargs(prm, u0[b], y0[b], y_r[b]), "update_cost_threshold": thr}, solve_on_init=False)
        rmse, u_ref, _ = eo.sim_control_loop(orc, lambda u: next(replay), T, d0[b], None, False)
        np.testing.assert_allclose(u_sys[:, b], u_ref, rtol=0, atol=2e-5)
        np.testing.assert_allclose(float(res.rmse[b]), rmse, rtol=1e-13, atol=0)
        np.testing.assert_array_equal(ctrl.u[b].cpu().numpy(), np.vstack([u0[b][T:], u_sys[:, b]]))  # the measurements always roll
        if thr is not None:
            assert abs(float(ctrl.tracking_cost[b]) - orc.tracking_cost) <= 1e-6 * max(1.0, orc.tracking_cost)
            np.testing.assert_array_equal(ctrl.u_data[b].cpu().numpy(), orc.u_data)  # the Hankel data only when the cost allows
            frozen += not np.array_equal(orc.u_data, orc.u)
        if alpha_reg_type == 1:
            np.testing.assert_allclose(ctrl.alpha_prev[b].cpu().numpy(), orc.alpha_prev, rtol=0, atol=2e-5)
    assert thr is None or frozen > 0


def test_early_stop_freezes_the_controller(build_lib):
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.controller_evaluation import (
        sim_nonlinear_dd_mpc_control_loop_batched,
    )

    B, T = 4, 8
    prm, env, ctrl, u0, y0, y_r = _setup(B, seed=5)
    # pretend the drones started 5 cm closer than they did and allow no increment: every drone whose
    # distance exceeds d0 at some step stops there (the measurement noise alone does not reach 5 cm)
    d_true = np.linalg.norm(env.get_pos(add_noise=False).cpu().numpy().astype(np.float64) - y_r, axis=1)
    d0 = d_true - 0.05
    res = sim_nonlinear_dd_mpc_control_loop_batched(env, ctrl, T, torch.from_numpy(d0), 0.0, False, record=True)
    y_sys = res.y_sys.cpu().numpy()
    fail_step = res.fail_step.cpu().numpy()
    for b in range(B):
        dist = np.linalg.norm(y_sys[:, b] - y_r[b], axis=1)
        over = np.nonzero(dist - d0[b] > 0.0)[0]
        expected = int(over[0]) if len(over) else -1
        assert fail_step[b] == expected
        if expected >= 0:
            assert bool(res.failed[b]) and np.isnan(float(res.rmse[b]))
            assert (y_sys[expected + 1 :, b] == 0).all()  # nothing recorded after the stop
            np.testing.assert_array_equal(ctrl.y[b].cpu().numpy()[-1], y_sys[expected, b])  # window frozen
    assert (fail_step >= 0).any()
