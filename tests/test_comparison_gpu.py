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


def test_comparison_run_against_the_published_reference_figure(build_lib):
    """The product's comparison run (GPU env + tracking kernel + fused policy MLP + DD-MPC solver, 16 replicas of each
    drone) against `docs/resources/controller_comparison_plot.png` of the reference, digitised into
    tests/golden/ref_comparison_plot_curves.npz (see tests/test_reference_plot_pin.py for the method):
      * tracking-controller drone and PPO drone: every pixel of the published curves within a few pixels of the mean
        flight (measured: max 2.1 / 4.9 px with seed 1; the tracking drones share the env - and its batch-global PID reset -
        with DD-MPC drones that may crash, hence the slightly wider bound than in the CPU test);
      * DD-MPC (third-party formulation, parity otherwise unpinned): the published MEAN curve (of an unknown, small number of
        runs) lies inside this implementation's mean +/- 2 sigma band at every pixel, and the `previous_equilibrium` reading
        of alpha_sr does not reproduce the figure (its runs diverge)."""
    import bench
    from test_reference_plot_pin import curve_stats
    from data_driven_quad_control_b200.comparison.utilities.batched_controller_sim import batched_controller_comparison
    from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
    from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type
    from data_driven_quad_control_b200.learning.policy import FusedActorMLP

    curves = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_comparison_plot_curves.npz"))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "policy_demo.npz"))
    w = {f"actor.{k}.{n}": torch.from_numpy(g[f"W{k}" + ("" if n == "weight" else "_b")]) for k in (0, 2, 4) for n in ("weight", "bias")}
    P = bench.MPC_DEFAULT
    ell = P["L"] + P["n"] + 1
    setpoints, steps, R = [[1.0, -1.0, 2.5], [0.0, 0.0, 1.5]], 175, 16
    at = EnvActionType.CTBR_FIXED_YAW

    def run(anchor):
        env_cfg, obs_cfg, _, cmd_cfg = get_cfgs()
        obs_cfg["obs_noise_std"] = 1e-4
        env_cfg.update(episode_length_s=100000, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0,
                       termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0)  # fmt: skip
        env = HoverEnv(3 * R, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device="cuda", action_type=at,
                       auto_target_updates=False, seed=6, materialize_buffers=True)  # fmt: skip
        kw = dict(n=P["n"], m=3, p=3, L=P["L"], N=P["N"], Q=np.kron(np.eye(ell), np.diag(P["Q"])), R=np.kron(np.eye(ell), np.diag(P["R"])),
                  S=np.diag(P["S"]), lamb_alpha=P["lamb_alpha"], lamb_sigma=P["lamb_sigma"], U=P["U"], Us=P["Us"],
                  alpha_reg_type=AlphaRegType.APPROXIMATED, lamb_alpha_s=P["lamb_alpha_s"], lamb_sigma_s=P["lamb_sigma_s"],
                  alpha_sr_anchor=anchor)  # fmt: skip
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)  # ground-plane watch of diverging DD-MPC runs
            traj = batched_controller_comparison(env, FusedActorMLP(w), kw, P["u_range"], P["init_hover_pos"], setpoints, steps,
                                                 np.random.default_rng(1))  # fmt: skip
        return np.stack(traj.system_outputs)  # (3 R, 350, 3)

    Y = run("optimal_reachable_equilibrium")
    for name, sl, lim in (("tracking", slice(0, R), (1.2, 2.5, 5.0)), ("rl", slice(R, 2 * R), (1.2, 3.6, 6.5))):
        st = curve_stats(curves, name, Y[sl].mean(0))
        for ax, (med, p90, mx, n) in st.items():
            assert med < lim[0] and p90 < lim[1] and mx < lim[2], (name, ax, med, p90, mx)
    t = np.arange(2 * steps) * 0.04

    def inside(Ym, k_sigma):
        ok = np.isfinite(Ym).all((1, 2)) & (np.abs(Ym).max((1, 2)) < 20)
        mean, std = Ym[ok].mean(0), Ym[ok].std(0)
        frac = []
        for i, ax in enumerate("xyz"):
            pts = curves[f"{ax}_dd_mpc"]
            m, s = np.interp(pts[:, 0], t, mean[:, i]), np.interp(pts[:, 0], t, std[:, i])
            frac.append(float((np.abs(pts[:, 1] - m) <= k_sigma * s + 2.0 / float(curves[f"{ax}_px_per_m"])).mean()))
        return int(ok.sum()), frac

    n_ok, frac = inside(Y[2 * R :], 2.0)
    assert n_ok == R and min(frac) > 0.98, (n_ok, frac)
    Yp = run("previous_equilibrium")[2 * R :]
    bad = int((~(np.isfinite(Yp).all((1, 2)) & (np.abs(Yp).max((1, 2)) < 20))).sum())
    spread = float(np.nanmean(Yp[:, 100:170, 0].std(0)))
    assert bad > 0 or spread > 3 * float(Y[2 * R :, 100:170, 0].std(0).mean()), (bad, spread)
