"""GPU parity of the fused env-step kernel against the numpy oracle (bit-exact contract)."""

import copy

import numpy as np
import pytest
import torch

from conftest import default_cfgs

pytestmark = pytest.mark.gpu


def _make_pair(N, at_name, auto=True, seed=5, env_over=None, obs_over=None, reward_cfg=None, materialize=True,
               independent=False):
    import env_oracle as eo
    from data_driven_quad_control_b200.controllers.ctbr.ctbr_controller import build_ctbr_matrices
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvCTBRControllerConfig, EnvDroneParams

    env_cfg, obs_cfg, rew_cfg, cmd_cfg, at = default_cfgs(at_name)
    if reward_cfg is not None:
        rew_cfg = reward_cfg
    env_cfg.update(env_over or {})
    obs_cfg.update(obs_over or {})
    cfgs_o = copy.deepcopy((env_cfg, obs_cfg, rew_cfg, cmd_cfg))
    env = HoverEnv(N, env_cfg, obs_cfg, rew_cfg, cmd_cfg, device="cuda", action_type=at, auto_target_updates=auto,
                   seed=seed, materialize_buffers=materialize, independent_envs=independent)
    mats = build_ctbr_matrices(EnvDroneParams.get())
    orc = eo.EnvOracle(N, *cfgs_o, action_type=at.value, auto_target_updates=auto, mixer=mats["mixer"].numpy(),
                       inv_sqrt_kf=float(mats["inv_sqrt_KF"]),
                       rate_pid_gains=EnvCTBRControllerConfig.get()["ctbr_controller_params"]["rate_pid_gains"],
                       rng=eo.PhiloxRng(seed), independent_envs=independent)
    return env, orc


def _actions(T, N, A, seed, gentle_frac=0.25):
    a = np.random.RandomState(seed).uniform(-1.3, 1.3, (T, N, A)).astype(np.float32)  # also exercises the clip
    a[:, : int(N * gentle_frac)] *= 0.03
    return a


def _compare_step(env, orc, t, exact=True, yaw_tol=0.0):
    o = env.obs_buf.cpu().numpy()
    assert np.array_equal(env.reset_buf.cpu().numpy(), orc.reset_buf), f"reset mask differs at step {t}"
    assert np.array_equal(env.extras["time_outs"].cpu().numpy(), orc.time_outs), f"time_outs differ at step {t}"
    assert np.array_equal(env.episode_length_buf.cpu().numpy(), orc.episode_length_buf)
    assert np.array_equal(env.hover_counter.cpu().numpy(), orc.hover_counter), f"hover counter differs at step {t}"
    if exact:
        assert np.array_equal(env.base_pos.cpu().numpy(), orc.base_pos), f"pos differs at step {t}"
        assert np.array_equal(env.base_quat.cpu().numpy(), orc.base_quat)
        assert np.array_equal(env.commands.cpu().numpy(), orc.commands), f"commands differ at step {t}"
        assert np.array_equal(o, orc.obs_buf), f"obs differs at step {t}: {np.abs(o - orc.obs_buf).max()}"
        assert np.array_equal(env.base_lin_vel.cpu().numpy(), orc.base_lin_vel)
        assert np.array_equal(env.base_ang_vel.cpu().numpy(), orc.base_ang_vel)
    r = env.rew_buf.cpu().numpy()
    if yaw_tol == 0.0:
        assert np.array_equal(r, orc.rew_buf), f"rew differs at step {t}: {np.abs(r - orc.rew_buf).max()}"
    else:
        np.testing.assert_allclose(r, orc.rew_buf, rtol=0, atol=yaw_tol)


@pytest.mark.parametrize("at_name", ["CTBR_FIXED_YAW", "CTBR", "ROTOR_RPMS"])
@pytest.mark.parametrize("auto", [True, False])
def test_random_rollout_bit_exact(build_lib, at_name, auto):
    """600 random-action steps x 96 envs, ~1.5k resets: trajectories, obs, rewards (yaw term: 1e-6),
    reset masks, time-outs, counters and resampled commands identical to the oracle."""
    N, T = 96, 600
    env, orc = _make_pair(N, at_name, auto)
    acts = _actions(T, N, env.num_actions, 1)
    env.reset()
    orc.reset()
    assert np.array_equal(env.commands.cpu().numpy(), orc.commands)
    # yaw reward (exp/atan2 from libm) only contributes when its scale is non-zero
    yaw_tol = 0.0 if at_name == "CTBR_FIXED_YAW" else 2e-6
    nres = 0
    for t in range(T):
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
        _compare_step(env, orc, t, yaw_tol=yaw_tol)
        nres += int(orc.reset_buf.sum())
        for k, v in orc.extras_episode.items():
            assert abs(float(env.extras["episode"][k]) - v) <= 1e-5 * max(1.0, abs(v)), (t, k)
    assert nres > 500
    for name, s in orc.episode_sums.items():
        np.testing.assert_allclose(env.episode_sums[name].cpu().numpy(), s, rtol=0, atol=1e-5 if yaw_tol else 0)


def test_hover_resample_and_fixup(build_lib):
    """Envs parked on their target exercise the hover counter, the Philox target resample and the
    A15/A16 stale-vs-fresh rel_pos rule on steps with and without a reset elsewhere."""
    N, T = 64, 120
    env, orc = _make_pair(N, "CTBR_FIXED_YAW", True)
    acts = _actions(T, N, 3, 2, gentle_frac=0.0)
    acts[:, :16] = 0.0  # zero action = hover thrust, zero rates
    acts[:20] = 0.0  # phase 1: nobody crashes, so the first hover events happen on reset-free steps
    env.reset()
    orc.reset()
    n_hover_events = 0
    n_fix_with_reset = n_fix_without_reset = 0
    for t in range(T):
        if t % 20 == 0:  # park the first 16 drones' targets on their current position
            idx = np.arange(16)
            tgt = orc.base_pos[idx].copy()
            orc.update_target_pos(idx, tgt)
            env.update_target_pos(torch.arange(16, device="cuda"), torch.from_numpy(tgt).cuda())
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
        _compare_step(env, orc, t)
        if orc.hover_resampled.any():
            n_hover_events += int(orc.hover_resampled.sum())
            if orc.reset_buf.any():
                n_fix_with_reset += 1
            else:
                n_fix_without_reset += 1
        np.testing.assert_array_equal(env.rel_pos.cpu().numpy(), orc.rel_pos)
        np.testing.assert_array_equal(env.last_rel_pos.cpu().numpy(), orc.last_rel_pos)
    assert n_hover_events >= 16
    assert n_fix_with_reset > 0 and n_fix_without_reset > 0


def test_lazy_buffers_match_materialized(build_lib):
    """materialize_buffers=False (the large-N mode) gives the same trajectory and the same
    body-frame velocities on demand."""
    N, T = 128, 60
    env_a, orc = _make_pair(N, "CTBR", True, materialize=False)
    acts = _actions(T, N, 4, 3)
    env_a.reset()
    orc.reset()
    for t in range(T):
        env_a.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
        _compare_step(env_a, orc, t, yaw_tol=2e-6)
    with pytest.raises(AttributeError):
        _ = env_a.base_euler


@pytest.mark.parametrize("N,T", [(4096, 300), (1 << 20, 3), ((1 << 20) - 77, 2)])
def test_baseline_sizes_bit_exact(build_lib, N, T):
    """BASELINE configs[1] (4096 envs) and configs[2] (2^20 envs per GPU, plus a ragged size whose last CTA is partly
    idle) on the kernel variant the bench times (CTBR_FIXED_YAW, lean buffers, auto target updates): every output of
    every env identical to the oracle, and the all-reduced rollout statistics equal to the oracle's sums."""
    env, orc = _make_pair(N, "CTBR_FIXED_YAW", True, materialize=False)
    acts = np.random.RandomState(11).uniform(-1.0, 1.0, (T, N, 3)).astype(np.float32)
    env.reset()
    orc.reset()
    n_reset, rew_sum = 0, 0.0
    for t in range(T):
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
        _compare_step(env, orc, t)
        n_reset += int(orc.reset_buf.sum())
        rew_sum += float(orc.rew_buf.astype(np.float64).sum())
    stats = env.rollout_stats()
    assert int(stats["resets"]) == n_reset
    assert int(stats["env_steps"]) == N * T
    assert abs(float(stats["reward_sum"]) - rew_sum) <= 2e-6 * max(1.0, abs(rew_sum))
    if N == 4096:
        assert n_reset > 1000  # the rollout exercises the reset / PID-flush / resample paths


@pytest.mark.parametrize("at_name,auto", [("CTBR_FIXED_YAW", False), ("CTBR", True), ("ROTOR_RPMS", True)])
def test_independent_envs_bit_exact(build_lib, at_name, auto):
    """independent_envs=True (every env its own num_envs = 1 instance: per-env PID reset and rel_pos refresh, the
    decoupled evaluation runs of the batched grid search) against the oracle in the same mode, incl. an
    external reset_idx in the middle of the rollout."""
    N, T = 64, 300
    env, orc = _make_pair(N, at_name, auto, independent=True)
    acts = _actions(T, N, env.num_actions, 7)
    env.reset()
    orc.reset()
    yaw_tol = 0.0 if at_name == "CTBR_FIXED_YAW" else 2e-6
    nres = 0
    for t in range(T):
        if t == 150:
            idx = [3, 17, 40]
            env.reset_idx(torch.tensor(idx, device="cuda"))
            orc.reset_idx(idx)
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
        _compare_step(env, orc, t, yaw_tol=yaw_tol)
        nres += int(orc.reset_buf.sum())
    assert nres > 200


def test_timeouts_and_episode_length_assignment(build_lib):
    """Zero actions hover forever: every env times out on step max_episode_length+1 (strict >)."""
    env, orc = _make_pair(8, "CTBR_FIXED_YAW", False, env_over={"episode_length_s": 0.4})
    env.reset()
    orc.reset()
    z = np.zeros((8, 3), np.float32)
    assert env.max_episode_length == 10
    for t in range(25):
        env.step(torch.from_numpy(z).cuda())
        orc.step(z)
        _compare_step(env, orc, t)
        if t == 10:
            assert env.reset_buf.all() and (env.extras["time_outs"] == 1.0).all()
    # rsl_rl-style assignment keeps working (init_at_random_ep_len)
    new = torch.randint(0, 10, (8,), device="cuda")
    env.episode_length_buf = new
    assert torch.equal(env.episode_length_buf, new.to(torch.int32))


def test_obs_noise_statistics(build_lib):
    env, orc = _make_pair(4096, "CTBR_FIXED_YAW", False, obs_over={"obs_noise_std": 0.05})
    env.reset()
    orc.reset()
    z = np.zeros((4096, 3), np.float32)
    env.step(torch.from_numpy(z).cuda())
    orc.step(z)
    o, oo = env.obs_buf.cpu().numpy(), orc.obs_buf
    # same Philox words, libm-level differences only
    np.testing.assert_allclose(o, oo, rtol=0, atol=5e-6)
    noise = o[:, 7:10]  # lin vel is ~0 while hovering
    assert abs(noise.std() - 0.05) < 0.004 and abs(noise.mean()) < 0.004
    assert np.array_equal(env.base_pos.cpu().numpy(), orc.base_pos)  # noise never touches the state


def test_actuator_noise(build_lib):
    env, orc = _make_pair(2048, "CTBR_FIXED_YAW", False, env_over={"actuator_noise_std": 50.0})
    env.reset()
    orc.reset()
    z = np.zeros((2048, 3), np.float32)
    env.step(torch.from_numpy(z).cuda())
    orc.step(z)
    np.testing.assert_allclose(env.rotor_rpms.cpu().numpy(), orc.last_rpm, rtol=0, atol=2e-2)
    np.testing.assert_allclose(env.base_pos.cpu().numpy(), orc.base_pos, rtol=0, atol=1e-6)


def test_state_save_restore_matches_oracle(build_lib):
    """get_current_state -> steps -> restore_from_state -> steps, against the oracle (incl. the
    body-frame-velocities-into-dofs quirk and the whole-batch PID state)."""
    import env_oracle as eo

    N = 6
    env, orc = _make_pair(N, "CTBR", False)
    acts = _actions(40, N, 4, 7, gentle_frac=1.0) * 10  # moderate actions, few resets
    env.reset()
    orc.reset()
    for t in range(8):
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
    idx = [2]
    s_env = env.get_current_state(torch.tensor(idx, device="cuda"))
    s_orc = orc.get_current_state(idx)
    for name in ("base_pos", "base_quat", "base_lin_vel", "base_ang_vel", "commands", "episode_length", "last_actions"):
        assert np.array_equal(getattr(s_env, name).cpu().numpy(), getattr(s_orc, name)), name
    assert np.array_equal(s_env.ctbr_controller_state.integral.cpu().numpy(), s_orc.pid_integral)
    assert np.array_equal(s_env.ctbr_controller_state.prev_error.cpu().numpy(), s_orc.pid_prev_error)
    for t in range(8, 16):
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
    env.restore_from_state(torch.tensor(idx, device="cuda"), s_env)
    orc.restore_from_state(idx, s_orc)
    again = env.get_current_state(torch.tensor(idx, device="cuda"))
    for name in ("base_pos", "base_quat", "base_lin_vel", "base_ang_vel", "episode_length", "last_actions"):
        assert torch.equal(getattr(again, name), getattr(s_env, name)), name  # reference test_drone_utilities.py
    # commands are NOT restored by the reference (hover_env.py:697-752 only derives rel_pos from the saved ones): the env
    # keeps the target it has now - a different one if it was reset since the save, as happens in this sequence
    assert np.array_equal(again.commands.cpu().numpy(), orc.get_current_state(idx).commands)
    assert np.array_equal(env.commands.cpu().numpy(), orc.commands)
    assert bool(env.reset_buf[2])
    for t in range(16, 40):
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
        _compare_step(env, orc, t, yaw_tol=2e-6)


@pytest.mark.parametrize("at_name", ["ROTOR_RPMS", "CTBR"])
def test_configs0_hover_to_target_500_steps(build_lib, at_name):
    """BASELINE configs[0]: single drone, SE(3) PID tracking controller (+ external CTBR controller
    in RPM mode), 500 steps from the reset pose to [0,0,1.5] (examples/drone_tracking_controller_example.py
    overrides).  The GPU stack (tracking kernel -> CTBR kernel -> fused env step) against the numpy
    oracles driven by the same loop: positions within 1e-4 relative over the 500 steps."""
    import controllers_oracle as co
    import env_oracle as eo
    from data_driven_quad_control_b200.controllers.ctbr.ctbr_controller import DroneCTBRController, build_ctbr_matrices
    from data_driven_quad_control_b200.controllers.tracking.tracking_controller_config import TrackingCtrlDroneState
    from data_driven_quad_control_b200.envs.hover_env_config import EnvCTBRControllerConfig, EnvDroneParams
    from data_driven_quad_control_b200.utilities.drone_tracking_controller import (
        create_drone_tracking_controller,
        tracking_env_action,
    )

    over = dict(episode_length_s=100, termination_if_close_to_ground=0.0, termination_if_x_greater_than=10.0,
                termination_if_y_greater_than=10.0, termination_if_z_greater_than=10.0)  # fmt: skip
    env, orc = _make_pair(1, at_name, False, env_over=over)
    env.reset()
    orc.reset()
    mats = build_ctbr_matrices(EnvDroneParams.get())
    gains = EnvCTBRControllerConfig.get()["ctbr_controller_params"]["rate_pid_gains"]
    trk = create_drone_tracking_controller(env)
    o_trk = co.TrackingOracle(0.027, [[2.0, 0.0, 2.2], [2.0, 0.0, 2.2], [18.0, 0.0, 6.0]], 5.0, 1, orc.step_dt)
    ctbr = o_ctbr = None
    if at_name == "ROTOR_RPMS":
        ctbr = DroneCTBRController(EnvDroneParams.get(), EnvCTBRControllerConfig.get(), env.step_dt, 1, env.device)
        o_ctbr = co.CTBROracle(mats["mixer"].numpy(), float(mats["inv_sqrt_KF"]), gains, 1, orc.step_dt)
    target = torch.tensor([[0.0, 0.0, 1.5]], device="cuda")
    q_sp = torch.tensor([[1.0, 0.0, 0.0, 0.0]], device="cuda")
    tgt = TrackingCtrlDroneState(X=target, Q=q_sp)
    lo, span = orc.act_lo, orc.act_span
    worst = 0.0
    for t in range(500):
        cur = TrackingCtrlDroneState(X=env.get_pos(), Q=env.get_quat())
        env.step(tracking_env_action(env, trk, tgt, cur, ctbr))
        cmd = o_trk.compute(orc.base_pos, orc.base_quat, target.cpu().numpy(), q_sp.cpu().numpy())
        if o_ctbr is not None:
            cmd = o_ctbr.compute(orc.base_ang_vel, cmd[:, 1:], cmd[:, 0])
        # linear_interpolate(x, lo, hi, -1, 1) in torch's op order: y_min + (x - x_min) * (y_max - y_min) / (x_max - x_min)
        orc.step((np.float32(-1.0) + (cmd - lo) * np.float32(2.0) / span).astype(np.float32))
        p_gpu, p_orc = env.base_pos.cpu().numpy(), orc.base_pos
        worst = max(worst, float(np.abs(p_gpu - p_orc).max() / max(1.0, np.abs(p_orc).max())))
        assert not orc.reset_buf.any() and not env.reset_buf.any()
    assert worst < 1e-4, worst
    assert np.abs(env.base_pos.cpu().numpy() - np.array([[0.0, 0.0, 1.5]])).max() < 5e-3  # it got there


def test_extras_episode_under_graph_replay(build_lib):
    """the extras["episode"] views must stay right when step() is replayed from a CUDA graph (the host
    names the ring row of every launch, DdqcEnvOut.ep_row, so nothing can drift).  A captured 8-step graph replayed
    3 times must leave every captured step's dict equal to what eager stepping produces for that step."""
    N, K, R = 256, 8, 3
    acts_np = _actions(K * (R + 1), N, 3, 11)
    acts = torch.from_numpy(acts_np).cuda()

    def fresh():
        env, _ = _make_pair(N, "CTBR_FIXED_YAW", True, materialize=False)
        env.reset()
        return env

    # eager reference: dict values of every step
    env_e = fresh()
    eager = []
    for t in range(K * (R + 1)):
        env_e.step(acts[t])
        eager.append({k: float(v) for k, v in env_e.extras["episode"].items()})
    pos_eager = env_e.base_pos.clone()

    env_g = fresh()
    static_a = torch.zeros((K, N, 3), device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    dicts = []
    with torch.cuda.stream(s):
        static_a.copy_(acts[:K])
        for t in range(K):  # warm-up = the first K steps, eagerly
            env_g.step(static_a[t])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for t in range(K):
                env_g.step(static_a[t])
                dicts.append(env_g.extras["episode"])  # views handed out during capture
        for r in range(R):
            static_a.copy_(acts[K * (r + 1) : K * (r + 2)])
            g.replay()
            s.synchronize()
            for t in range(K):
                want = eager[K * (r + 1) + t]
                got = {k: float(v) for k, v in dicts[t].items()}
                assert got == want, (r, t, got, want)
    torch.cuda.synchronize()
    assert torch.equal(env_g.base_pos, pos_eager)
    # eager stepping after the replays continues on fresh rows and persists the last values when nobody resets
    env_g.step(torch.zeros((N, 3), device="cuda"))
    assert set(env_g.extras["episode"]) == set(eager[-1])


def test_restore_with_empty_index_consumes_no_ring_row(build_lib):
    env, _ = _make_pair(32, "CTBR_FIXED_YAW", True)
    env.reset()
    for _ in range(3):
        env.step(torch.zeros((32, 3), device="cuda"))
    st = env.get_current_state(torch.arange(0, device="cuda"))
    row = env._ep_row
    env.restore_from_state(torch.arange(0, device="cuda"), st)
    assert env._ep_row == row
    env.step(torch.zeros((32, 3), device="cuda"))
    raw = env._ctl.cpu().numpy().tobytes()
    from data_driven_quad_control_b200 import _lib

    assert _lib.EnvCtl.from_buffer_copy(raw).ep_row == env._ep_row  # host and device agree on the current row


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_env_on_a_non_current_device(build_lib):
    """objects built on cuda:1 while cuda:0 is current launch on their own device."""
    from data_driven_quad_control_b200.learning.policy import FusedActorMLP

    assert torch.cuda.current_device() == 0
    env1, orc = _make_pair(64, "CTBR_FIXED_YAW", True)
    env_cfg, obs_cfg, rew_cfg, cmd_cfg, at = default_cfgs("CTBR_FIXED_YAW")
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv

    env = HoverEnv(64, env_cfg, obs_cfg, rew_cfg, cmd_cfg, device="cuda:1", action_type=at, seed=5, materialize_buffers=True)
    acts = _actions(20, 64, 3, 4)
    env.reset()
    orc.reset()
    for t in range(20):
        env.step(torch.from_numpy(acts[t]).to("cuda:1"))
        orc.step(acts[t])
        assert np.array_equal(env.obs_buf.cpu().numpy(), orc.obs_buf)
    pol = FusedActorMLP.random_init(device="cuda:1")
    a1 = pol.act_inference(env.obs_buf)
    a0 = FusedActorMLP.random_init(device="cuda:0").act_inference(env.obs_buf.to("cuda:0"))
    assert torch.equal(a1.cpu(), a0.cpu())
    assert torch.cuda.current_device() == 0


def test_ground_plane_watch_counts_what_the_reference_floor_would_stop(build_lib):
    """No ground plane is modelled (reference: hover_env.py:199).  With the low-altitude termination disabled - the
    tracking example, the DD-MPC evaluation and grid-search configurations - a drone with zero thrust falls through
    z = 0; the kernel counts those env-steps exactly like the oracle and `check_ground_plane` warns.  With the default
    termination (z < 0.1 resets the env) nothing is counted."""
    N, T = 64, 40
    over = dict(termination_if_close_to_ground=0.0, termination_if_z_greater_than=50.0, termination_if_lin_vel_greater_than=1e3)
    env, orc = _make_pair(N, "CTBR_FIXED_YAW", False, env_over=over)
    acts = np.zeros((T, N, 3), np.float32)
    acts[:, : N // 2, 0] = -1.0  # zero thrust: free fall from z = 1 reaches the floor after ~0.45 s = 12 steps
    env.reset()
    orc.reset()
    for t in range(T):
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
        _compare_step(env, orc, t)
    assert orc.below_ground_env_steps > 0
    assert int(env.rollout_stats()["below_ground_env_steps"]) == orc.below_ground_env_steps
    with pytest.warns(RuntimeWarning, match="below the ground plane"):
        assert env.check_ground_plane("test") == orc.below_ground_env_steps
    env2, _ = _make_pair(N, "CTBR_FIXED_YAW", False)
    env2.reset()
    for t in range(T):
        env2.step(torch.from_numpy(acts[t]).cuda())
    assert env2.check_ground_plane() == 0
