"""Developer probe: the product's three-way comparison run (GPU) against the digitised reference figure
(tests/golden/ref_comparison_plot_curves.npz): pixel distances of the tracking / policy / DD-MPC-mean curves, for both
alpha_sr anchors and a few replica counts.  Prints one JSON line per configuration."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench
from test_reference_plot_pin import curve_stats, pixel_distance
from data_driven_quad_control_b200.comparison.utilities.batched_controller_sim import batched_controller_comparison
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType
from data_driven_quad_control_b200.envs.hover_env import HoverEnv
from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type
from data_driven_quad_control_b200.learning.policy import FusedActorMLP

curves = np.load(os.path.join(ROOT, "tests", "golden", "ref_comparison_plot_curves.npz"))
g = np.load(os.path.join(ROOT, "tests", "golden", "policy_demo.npz"))
w = {f"actor.{k}.{n}": torch.from_numpy(g[f"W{k}" + ("" if n == "weight" else "_b")]) for k in (0, 2, 4) for n in ("weight", "bias")}
P = bench.MPC_DEFAULT
ell = P["L"] + P["n"] + 1
setpoints, steps = [[1.0, -1.0, 2.5], [0.0, 0.0, 1.5]], 175
for anchor in ("optimal_reachable_equilibrium", "previous_equilibrium"):
    for R, seed in ((32, 0), (32, 1)):
        env_cfg, obs_cfg, _, cmd_cfg = get_cfgs()
        obs_cfg["obs_noise_std"] = 1e-4
        env_cfg.update(episode_length_s=100000, termination_if_close_to_ground=0.0, termination_if_x_greater_than=100.0,
                       termination_if_y_greater_than=100.0, termination_if_z_greater_than=10.0)
        at = EnvActionType.CTBR_FIXED_YAW
        env = HoverEnv(3 * R, env_cfg, obs_cfg, get_reward_cfg_by_action_type(at), cmd_cfg, device="cuda", action_type=at,
                       auto_target_updates=False, seed=5 + seed, materialize_buffers=True)
        kw = dict(n=P["n"], m=3, p=3, L=P["L"], N=P["N"], Q=np.kron(np.eye(ell), np.diag(P["Q"])), R=np.kron(np.eye(ell), np.diag(P["R"])),
                  S=np.diag(P["S"]), lamb_alpha=P["lamb_alpha"], lamb_sigma=P["lamb_sigma"], U=P["U"], Us=P["Us"],
                  alpha_reg_type=AlphaRegType.APPROXIMATED, lamb_alpha_s=P["lamb_alpha_s"], lamb_sigma_s=P["lamb_sigma_s"], alpha_sr_anchor=anchor)
        traj = batched_controller_comparison(env, FusedActorMLP(w), kw, P["u_range"], P["init_hover_pos"], setpoints, steps, np.random.default_rng(seed))
        Y = np.stack(traj.system_outputs)  # (3R, 350, 3)
        out = {"anchor": anchor, "R": R, "seed": seed}
        for name, sl in (("tracking", slice(0, R)), ("rl", slice(R, 2 * R)), ("dd_mpc", slice(2 * R, 3 * R))):
            ok = np.isfinite(Y[sl]).all((1, 2)) & (np.abs(Y[sl]).max((1, 2)) < 20)
            mean = Y[sl][ok].mean(0)
            st = curve_stats(curves, name, mean)
            out[name] = {ax: [round(v, 2) for v in st[ax][:3]] for ax in "xyz"}
            out[name]["runs_ok"] = int(ok.sum())
            if name == "dd_mpc":
                std = Y[sl][ok].std(0)
                out[name]["peak_x"] = [round(float(mean[:175, 0].max()), 3), round(float(np.argmax(mean[:175, 0]) * 0.04), 2)]
                out[name]["min_y"] = [round(float(mean[:175, 1].min()), 3), round(float(np.argmin(mean[:175, 1]) * 0.04), 2)]
                out[name]["peak_z"] = [round(float(mean[:175, 2].max()), 3), round(float(np.argmax(mean[:175, 2]) * 0.04), 2)]
                out[name]["std_x_at_2s"] = round(float(std[50, 0]), 3)
                np.save(os.path.join(ROOT, "gpurun_out", f"ddmpc_mean_{anchor[:4]}_{seed}.npy"), np.concatenate([mean, std], 1))
        print(json.dumps(out), flush=True)
