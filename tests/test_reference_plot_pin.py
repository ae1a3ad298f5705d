"""The restated physics against the ONE output of the real reference stack that ships with the reference:
`docs/resources/controller_comparison_plot.png` (Genesis + the reference's controllers, drawn by
`comparison/controller_comparison_plot.py:60-186` from runs of `comparison/controller_comparison.py`), digitised by
`oracle/digitize_reference_plot.py` into `tests/golden/ref_comparison_plot_curves.npz`.

The scenario of `configs/comparison/controller_comparison_config.yaml` (hover at [0, 0, 1.5]; setpoints [1, -1, 2.5] and
[0, 0, 1.5], 175 control steps each; CTBR_FIXED_YAW actions with latency) is flown on the oracle env with
  * the SE(3) tracking controller (reference gains, `configs/controllers/tracking/tracking_controller_params.yaml`) and
  * the reference's trained PPO policy (`models/demo_ctbr_fixed_yaw/model_1500.pt`, weights in tests/golden/policy_demo.npz),
and every digitised pixel of the published red / green curves must lie within a few pixels of the simulated trajectory
(distance measured in the figure's own pixel metric: 21.3 px per second, 148-173 px per metre; the lines are 2 px wide).
The CUDA env is bit-identical to the oracle env (tests/test_env_parity_gpu.py), so this pins the product's physics too.
"""

import copy
import os

import numpy as np
import pytest

import ddmpc_oracle as mo
import env_oracle as eo
import policy_oracle as po
from controllers_oracle import TrackingOracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DT, STEPS = 0.04, 175
SETPOINTS = ([1.0, -1.0, 2.5], [0.0, 0.0, 1.5])


def comparison_env():
    """The env of comparison/controller_comparison.py:186-229 with one drone (no observation noise: 0.3 mm)."""
    env_cfg = dict(
        dt=0.01, decimation=4, simulate_action_latency=True, clip_actions=1.0, actuator_noise_std=0.0,
        termination_if_roll_greater_than=180, termination_if_pitch_greater_than=180, termination_if_close_to_ground=0.0,
        termination_if_x_greater_than=100.0, termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0,
        termination_if_ang_vel_greater_than=20, termination_if_lin_vel_greater_than=20, base_init_pos=[0.0, 0.0, 1.0],
        base_init_quat=[1.0, 0.0, 0.0, 0.0], episode_length_s=100, at_target_threshold=0.1, min_hover_time_s=0.5,
        resampling_time_s=3.0,
    )  # fmt: skip
    obs_cfg = dict(obs_scales=dict(rel_pos=1 / 3.0, lin_vel=1 / 3.0, ang_vel=1 / 3.14159), obs_noise_std=0.0)
    rew = dict(yaw_lambda=-10.0, reward_scales=dict(target=10.0, closeness=1.5, hover_time=0.01, smooth=-1e-4, yaw=0.01,
                                                     angular=-2e-4, crash=-10.0))  # fmt: skip
    cmd = dict(num_commands=3, pos_x_range=[-1.0, 1.0], pos_y_range=[-1.0, 1.0], pos_z_range=[1.0, 3.0])
    mixer, isk = mo.ctbr_mixer()
    env = eo.EnvOracle(1, env_cfg, obs_cfg, copy.deepcopy(rew), cmd, action_type=eo.CTBR_FIXED_YAW, auto_target_updates=False,
                       mixer=mixer, inv_sqrt_kf=isk, rng=eo.PhiloxRng(1))  # fmt: skip
    env.reset()
    return env


def fly(which: str, with_inputs: bool = False) -> np.ndarray:
    """(350, 3) true positions logged after every control step, as parallel_controller_sim.py:339-340 records them
    (`with_inputs`: (350, 6), the applied thrust [N] / roll / pitch rate commands [rad/s] appended, :337, :425-431)."""
    env = comparison_env()
    ctrl = TrackingOracle(0.027, [[2.0, 0.0, 2.2], [2.0, 0.0, 2.2], [18.0, 0.0, 6.0]], 5.0, 1, env.step_dt)
    q_sp = np.array([[1.0, 0.0, 0.0, 0.0]], np.float32)
    lo, span = env.act_lo.astype(np.float64), env.act_span.astype(np.float64)

    def track(target, steps, log):
        tgt = np.array([target], np.float32)
        env.update_target_pos([0], tgt)
        for _ in range(steps):
            c = ctrl.compute(env.base_pos.astype(np.float32), env.base_quat, tgt, q_sp)[:, :3].astype(np.float64)
            a = np.clip(-1.0 + (c - lo) * 2.0 / span, -1, 1)
            env.step(a.astype(np.float32))
            if log is not None:
                log.append(np.concatenate([env.base_pos[0], (lo + (a + 1.0) * span / 2.0)[0]]))

    # stabilisation at the initial hover position + the 350 steps of the DD-MPC drone's data collection, during which the
    # other drones are held by the stabilising (tracking) controller (controller_comparison.py:286-330)
    track([0.0, 0.0, 1.5], 500, None)
    log = []
    if which == "tracking":
        for sp in SETPOINTS:
            track(sp, STEPS, log)
    else:
        g = np.load(os.path.join(GOLDEN, "policy_demo.npz"))
        W = (g["W0"], g["W0_b"], g["W2"], g["W2_b"], g["W4"], g["W4_b"])
        obs = env.obs_buf.copy()  # the worker acts on the observation of the previous step (rl_worker.py)
        for sp in SETPOINTS:
            env.update_target_pos([0], np.array([sp], np.float32))
            for _ in range(STEPS):
                a = np.clip(po.actor_forward(obs, *W), -1, 1)
                env.step(a.astype(np.float32))
                obs = env.obs_buf.copy()
                log.append(np.concatenate([env.base_pos[0], (lo + (a.astype(np.float64) + 1.0) * span / 2.0)[0]]))
    log = np.array(log)
    return log if with_inputs else log[:, :3]


def pixel_distance(points: np.ndarray, values: np.ndarray, px_per_s: float, px_per_unit: float) -> np.ndarray:
    """Distance, in figure pixels, from every digitised point (t, value) to the polyline of a simulated trajectory
    (sample k drawn at t = k * 0.04 s, like `ax.plot(range(T), ...)` with the axis relabelled in seconds)."""
    t = np.arange(len(values)) * DT
    tf = np.linspace(t[0], t[-1], len(t) * 20)
    P = np.stack([tf * px_per_s, np.interp(tf, t, values) * px_per_unit], 1)
    Q = np.stack([points[:, 0] * px_per_s, points[:, 1] * px_per_unit], 1)
    return np.sqrt(((Q[:, None, :] - P[None, :, :]) ** 2).sum(2)).min(1)


def curve_stats(curves, key: str, pos: np.ndarray) -> dict:
    out = {}
    for i, ax in enumerate("xyz"):
        d = pixel_distance(curves[f"{ax}_{key}"], pos[:, i], float(curves[f"{ax}_px_per_s"]), float(curves[f"{ax}_px_per_m"]))
        out[ax] = (float(np.median(d)), float(np.percentile(d, 90)), float(d.max()), len(d))
    return out


@pytest.fixture(scope="module")
def curves():
    return np.load(os.path.join(GOLDEN, "ref_comparison_plot_curves.npz"))


def test_digitised_curves_are_sane(curves):
    """Enough pixels of every line, inside the axes, and the curves end on the second setpoint to a pixel."""
    for ax, sp in zip("xyz", SETPOINTS[1]):
        for key in ("tracking", "rl", "dd_mpc"):
            pts = curves[f"{ax}_{key}"]
            assert len(pts) > 250 and pts[:, 0].min() >= 0 and pts[:, 0].max() <= 14.0
            tail = pts[pts[:, 0] > 12.0, 1]  # (the red line is hidden under the other two at the end)
            assert key == "tracking" or abs(np.median(tail) - sp) < 0.04, (ax, key, np.median(tail))


def test_tracking_controller_flight_matches_the_published_genesis_run(curves):
    """Deterministic: the tracking controller's flight depends on the physics only.  Measured: median 0.6 px, 90th
    percentile <= 1.43 px, max 2.0 / 2.1 / 3.1 px (x / y / z) - inside the drawn line everywhere, including the violent
    second transition (the derivative kick of the 1 m step down commands zero thrust and a momentary inversion, y
    overshoots to +0.17 m).  With the velocity-first substep order used until round 2: max 3.6 / 9.8 / 4.7 px (overshoot
    +0.10 m)."""
    st = curve_stats(curves, "tracking", fly("tracking"))
    for ax, (med, p90, mx, n) in st.items():
        assert med < 1.0 and p90 < 2.0 and mx < 4.0, (ax, med, p90, mx, n)


def test_ppo_policy_flight_matches_the_published_genesis_run(curves):
    """The reference's trained policy closes the loop through the full 16-dimensional observation (body-frame velocities,
    rates, attitude, last action), so its flight probes the fast attitude dynamics.  Measured: median <= 0.72 px, 90th
    percentile 1.6 / 3.0 / 1.05 px, max 3.8 / 4.9 / 2.5 px, and a hover as steady as the reference's (std 0.2 mm over
    t = 4 .. 6.8 s; the published curve is flat to the pixel).  Velocity-first order: max 7.7 / 12.6 / 10.2 px and a 20 mm
    limit cycle in hover."""
    pos = fly("rl")
    st = curve_stats(curves, "rl", pos)
    for ax, (med, p90, mx, n) in st.items():
        assert med < 1.0 and p90 < 3.6 and mx < 6.0, (ax, med, p90, mx, n)
    assert pos[100:170].std(0).max() < 2e-3
    # same steady-state offsets as the reference's run (the policy is a proportional-type controller): read off the figure
    ref_hold = [np.median(curves[f"{ax}_rl"][(curves[f"{ax}_rl"][:, 0] > 4.0) & (curves[f"{ax}_rl"][:, 0] < 6.8), 1]) for ax in "xyz"]
    assert np.abs(pos[100:170].mean(0) - ref_hold).max() < 0.012, (pos[100:170].mean(0), ref_hold)


def test_ppo_policy_commands_match_the_published_genesis_run(curves):
    """The "Control Inputs" row of the same figure: total thrust and roll / pitch rate commands of the PPO drone (the
    green lines; 462 px per newton, 8.2 px per rad/s).  The commands ring after each setpoint change (lightly damped
    attitude loop), which makes them the most sensitive probe of the fast dynamics the figure offers.  Measured: median
    0.9 - 1.1 px, 90th percentile 2.0 / 2.2 / 2.1 px (velocity-first order: thrust 90th percentile 4.3 px)."""
    log = fly("rl", with_inputs=True)
    for i, ax in enumerate(("thrust", "roll", "pitch")):
        d = pixel_distance(curves[f"{ax}_rl"], log[:, 3 + i], float(curves[f"{ax}_px_per_s"]), float(curves[f"{ax}_px_per_m"]))
        assert np.median(d) < 1.5 and np.percentile(d, 90) < 3.0, (ax, np.median(d), np.percentile(d, 90))


@pytest.mark.reference
def test_fixture_is_what_the_digitiser_reads_from_the_reference_figure(tmp_path, curves, monkeypatch):
    """Build container only: `oracle/digitize_reference_plot.py` run on /root/reference/docs/resources/
    controller_comparison_plot.png reproduces the committed fixture exactly."""
    import digitize_reference_plot as dg

    out = tmp_path / "curves.npz"
    monkeypatch.setattr(dg, "OUT", str(out))
    dg.main()
    fresh = np.load(out)
    assert sorted(fresh.files) == sorted(curves.files)
    for k in fresh.files:
        assert np.array_equal(fresh[k], curves[k]), k
