"""GPU re-statement of the reference's integration tests against the B200 classes
(reference tests/controller_tests/test_drone_ctbr_controller.py:17-104,
tests/controller_tests/test_drone_tracking_controller.py:23-82,
tests/utility_tests/test_drone_utilities.py:18-76, tests/env_tests/test_hover_env.py:14-60),
plus parity of the stand-alone controller kernels with the reference classes' golden outputs."""

import os
from dataclasses import asdict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _dummy_cfgs():
    """tests/conftest.py:24-87 of the reference."""
    env_cfg = dict(
        dt=0.01, decimation=4, simulate_action_latency=True, clip_actions=1.0, actuator_noise_std=0.0,
        termination_if_roll_greater_than=180, termination_if_pitch_greater_than=180,
        termination_if_close_to_ground=0.1, termination_if_x_greater_than=3.0, termination_if_y_greater_than=3.0,
        termination_if_z_greater_than=2.0, termination_if_ang_vel_greater_than=12,
        termination_if_lin_vel_greater_than=20, base_init_pos=[0.0, 0.0, 1.0], base_init_quat=[1.0, 0.0, 0.0, 0.0],
        episode_length_s=15.0, at_target_threshold=0.1, min_hover_time_s=0.01, resampling_time_s=3.0,
        visualize_target=False, visualize_camera=False, max_visualize_FPS=100,
    )  # fmt: skip
    obs_cfg = dict(obs_scales=dict(rel_pos=1.0, lin_vel=1.0, ang_vel=1.0), obs_noise_std=0.0)
    rew_cfg = dict(yaw_lambda=-1.0, reward_scales=dict(target=1.0, closeness=1.0, hover_time=1.0, smooth=1.0, yaw=1.0,
                                                        angular=1.0, crash=1.0))  # fmt: skip
    cmd_cfg = dict(num_commands=3, pos_x_range=(-1.0, 1.0), pos_y_range=(-1.0, 1.0), pos_z_range=(0.5, 2.0))
    return env_cfg, obs_cfg, rew_cfg, cmd_cfg


def test_hover_env_loop(build_lib):
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType

    cfgs = _dummy_cfgs()
    env = HoverEnv(2, *cfgs, show_viewer=False, device="cuda", action_type=EnvActionType.CTBR_FIXED_YAW)
    assert cfgs[2]["reward_scales"]["target"] == pytest.approx(0.04)  # scaled in the caller's dict (A23)
    obs, _ = env.reset()
    assert obs.shape == (2, env.num_obs)
    for _ in range(5):
        obs, reward, done, info = env.step(torch.zeros((2, env.num_actions), device=env.device))
        assert obs.shape == (2, env.num_obs) and reward.shape == (2,) and done.shape == (2,) and isinstance(info, dict)
        assert not torch.isnan(reward).any()
    for k in cfgs[2]["reward_scales"]:
        assert f"rew_{k}" in env.extras["episode"]
    assert env.extras["time_outs"].dtype == torch.float32 and env.reset_buf.dtype == torch.bool
    o2, ex2 = env.get_observations()
    assert o2 is env.obs_buf and ex2 is env.extras


def test_drone_ctbr_controller_behaviour(build_lib):
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, EnvDroneParams
    from data_driven_quad_control_b200.utilities.math_utils import linear_interpolate

    env = HoverEnv(2, *_dummy_cfgs(), device="cuda", action_type=EnvActionType.CTBR)
    env.reset()

    def run(setpoint, steps):
        sp = torch.tensor([setpoint], dtype=torch.float, device=env.device).expand(2, -1)
        a = linear_interpolate(x=sp, x_min=env.action_bounds[:, 0], x_max=env.action_bounds[:, 1], y_min=-1, y_max=1)
        for _ in range(steps):
            env.step(a)
        return sp

    run([EnvDroneParams.WEIGHT, 0.0, 0.0, 0.0], 10)
    assert torch.all(torch.abs(env.base_ang_vel) < 0.1), "Hover control fails"
    sp = run([EnvDroneParams.WEIGHT, 0.01, -0.01, 1.0], 20)
    torch.testing.assert_close(env.base_ang_vel, sp[:, 1:], rtol=0.0, atol=1e-3)


@pytest.mark.parametrize("action_type_name", ["ROTOR_RPMS", "CTBR", "CTBR_FIXED_YAW"])
def test_tracking_controller_reaches_targets(build_lib, action_type_name):
    from data_driven_quad_control_b200.controllers.ctbr.ctbr_controller import DroneCTBRController
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import (
        EnvActionType,
        EnvCTBRControllerConfig,
        EnvDroneParams,
    )
    from data_driven_quad_control_b200.utilities.drone_tracking_controller import (
        create_drone_tracking_controller,
        hover_at_target,
    )

    at = EnvActionType[action_type_name]
    cfgs = _dummy_cfgs()
    cfgs[0].update(episode_length_s=100, termination_if_close_to_ground=0.0, termination_if_x_greater_than=10.0,
                   termination_if_y_greater_than=10.0, termination_if_z_greater_than=10.0)  # fmt: skip
    env = HoverEnv(2, *cfgs, device="cuda", action_type=at, auto_target_updates=False)
    env.reset()
    trk = create_drone_tracking_controller(env)
    ctbr = None
    if at == EnvActionType.ROTOR_RPMS:
        ctbr = DroneCTBRController(EnvDroneParams.get(), EnvCTBRControllerConfig.get(), env.step_dt, 2, env.device)
    target = torch.tensor([[0.0, 0.0, 1.5], [0.5, -0.5, 1.2]], device=env.device)
    steps = hover_at_target(env, trk, target, torch.zeros(2, device=env.device), min_at_target_steps=10,
                            error_threshold=5e-2, ctbr_controller=ctbr, max_steps=500)  # fmt: skip
    assert steps < 500, "tracking controller did not settle"
    assert torch.abs(env.base_pos - target).max() < 5e-2
    assert not env.reset_buf.any()


def test_hover_env_utilities_save_restore(build_lib):
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType
    from data_driven_quad_control_b200.utilities.drone_environment import (
        create_env,
        get_current_env_state,
        restore_env_from_state,
        update_env_target_pos,
    )

    env = create_env(2, *_dummy_cfgs(), show_viewer=False, device="cuda", action_type=EnvActionType.CTBR_FIXED_YAW)
    obs, _ = env.reset()
    assert obs.shape == (2, env.num_obs)
    update_env_target_pos(env=env, env_idx=0, target_pos=torch.tensor([0.0, 0.0, 1.5], device=env.device))
    assert env.commands[0].tolist() == [0.0, 0.0, 1.5]
    s0 = get_current_env_state(env=env, env_idx=0)
    for _ in range(5):
        env.step(torch.zeros((2, env.num_actions), device=env.device))
    restore_env_from_state(env=env, env_idx=0, saved_state=s0)
    s1 = get_current_env_state(env=env, env_idx=0)

    def same(a, b):
        for k in a:
            if isinstance(a[k], dict):
                if not same(a[k], b[k]):
                    return False
            elif not torch.equal(a[k], b[k]):
                return False
        return a.keys() == b.keys()

    assert same(asdict(s0), asdict(s1)), "Env states differ after restore."


def test_standalone_controllers_match_reference_golden(build_lib):
    """ddqc_ctbr_compute / ddqc_tracking_compute vs outputs of the reference classes (controllers.npz)
    and, bit for bit, vs the numpy controller oracles; strided (transposed) inputs included."""
    import controllers_oracle as co

    from data_driven_quad_control_b200.controllers.ctbr.ctbr_controller import DroneCTBRController
    from data_driven_quad_control_b200.controllers.tracking.tracking_controller import DroneTrackingController
    from data_driven_quad_control_b200.controllers.tracking.tracking_controller_config import TrackingCtrlDroneState
    from data_driven_quad_control_b200.envs.hover_env_config import EnvCTBRControllerConfig, EnvDroneParams
    from data_driven_quad_control_b200.utilities.config_utils import load_yaml_config
    from data_driven_quad_control_b200.utilities.drone_tracking_controller import TRACKING_CONTROLLER_CONFIG_PATH
    from data_driven_quad_control_b200.utilities.vectorized_pid_controller import VectorizedPIDController

    g = np.load(os.path.join(GOLD, "controllers.npz"))
    T, N = g["rate_meas"].shape[:2]
    dev = torch.device("cuda")
    ctbr = DroneCTBRController(EnvDroneParams.get(), EnvCTBRControllerConfig.get(), dt=0.04, num_envs=N, device=dev)
    # the host-built mixer equals the reference's (same torch CPU calls)
    np.testing.assert_array_equal(ctbr.mixer.cpu().numpy(), g["mixer"])
    trk = DroneTrackingController(EnvDroneParams.MASS, load_yaml_config(TRACKING_CONTROLLER_CONFIG_PATH), 0.04, N, dev)
    o_ctbr = co.CTBROracle(g["mixer"], g["inv_sqrt_kf"], [[10.0, 0.0, 0.1]] * 3, N, 0.04)
    o_trk = co.TrackingOracle(0.027, [[2.0, 0.0, 2.2], [2.0, 0.0, 2.2], [18.0, 0.0, 6.0]], 5.0, N, 0.04)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    for t in range(T):
        meas = cu(g["rate_meas"][t].T).T  # (N,3) view of (3,N) storage: exercises the stride path
        rpm = ctbr.compute(meas, cu(g["rate_sp"][t]), cu(g["thrust_sp"][t])).cpu().numpy()
        np.testing.assert_array_equal(rpm, o_ctbr.compute(g["rate_meas"][t], g["rate_sp"][t], g["thrust_sp"][t]))
        np.testing.assert_allclose(rpm**2, g["rpm"][t] ** 2, rtol=1e-5, atol=300.0)
        cmd = trk.compute(
            state_measurement=TrackingCtrlDroneState(X=cu(g["X"][t]), Q=cu(g["Q"][t].T).T),
            state_setpoint=TrackingCtrlDroneState(X=cu(g["X_sp"]), Q=cu(g["Q_sp"])),
        ).cpu().numpy()
        np.testing.assert_allclose(cmd, o_trk.compute(g["X"][t], g["Q"][t], g["X_sp"], g["Q_sp"]), rtol=0, atol=2e-6)
        np.testing.assert_allclose(cmd, g["ctbr_cmd"][t], rtol=1e-5, atol=2e-5)
    # PID state API (vectorized_pid_controller.py:128-158)
    st = ctbr.get_state()
    assert st.integral.shape == (N, 3) and torch.equal(st.prev_error, ctbr.rate_pid_controller.prev_error)
    ctbr.reset()
    assert not ctbr.rate_pid_controller.prev_error.any()
    ctbr.load_state(st)
    assert torch.equal(ctbr.rate_pid_controller.prev_error, st.prev_error)
    pid = VectorizedPIDController(torch.tensor([[1.0, 0.5, 0.1], [2.0, 0.0, 0.0]]), 3, 0.1, dev)
    out = pid.compute(torch.ones(3, 2, device=dev), torch.zeros(3, 2, device=dev))
    torch.testing.assert_close(out.cpu(), torch.tensor([[1.0 + 0.05 + 1.0, 2.0]] * 3))
    trk.reset()
    assert not trk.controller.integral.any()
