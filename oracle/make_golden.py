"""TEST INFRASTRUCTURE ONLY -- regenerates the fixtures under tests/golden/.

Run in the build container (needs the read-only reference tree):

    python oracle/make_golden.py

What each file pins, and from which source of truth:

  controllers.npz   outputs of the REFERENCE classes DroneCTBRController / DroneTrackingController /
                    VectorizedPIDController imported from /root/reference/src (torch CPU), plus the
                    mixer they build.  Pins oracle/controllers_oracle.py and the CUDA controllers.
  env_<type>.npz    rollouts of the REFERENCE `envs/hover_env.py`, imported unchanged, running on
                    oracle/genesis_shim (torch CPU, seeded global generator): actions, obs, rewards,
                    reset masks, time-outs, commands, hover counters, extras.  Pins the env LOGIC of
                    oracle/env_oracle.py (reset indices exact; floats to tolerance because the
                    reference's mixer matmul goes through BLAS).
  ddmpc_*.npz       solutions of the literal DD-MPC QP by oracle/ddmpc_oracle.py with their KKT
                    residuals (a solver-independent optimality certificate).  Parity against the
                    third-party cvxpy implementation itself is UNPINNED (not installable offline).
"""

from __future__ import annotations

import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path[:0] = [HERE, os.path.join(HERE, "genesis_shim"), "/root/reference/src"]


def make_controllers():
    import torch
    from data_driven_quad_control.controllers.ctbr.ctbr_controller import DroneCTBRController
    from data_driven_quad_control.controllers.tracking.tracking_controller import DroneTrackingController
    from data_driven_quad_control.controllers.tracking.tracking_controller_config import TrackingCtrlDroneState
    from data_driven_quad_control.envs.hover_env_config import EnvCTBRControllerConfig, EnvDroneParams
    from data_driven_quad_control.utilities.config_utils import load_yaml_config

    N, T = 64, 5
    rng = np.random.default_rng(0)
    ctbr = DroneCTBRController(EnvDroneParams.get(), EnvCTBRControllerConfig.get(), dt=0.04, num_envs=N, device="cpu")
    meas = rng.uniform(-3, 3, (T, N, 3)).astype(np.float32)
    sp = rng.uniform(-3, 3, (T, N, 3)).astype(np.float32)
    thrust = rng.uniform(0.0, 0.53, (T, N)).astype(np.float32)
    # the two hand-checked cases of SURVEY.md 8(c)
    meas[0, 0], sp[0, 0], thrust[0, 0] = 0.0, 0.0, EnvDroneParams.WEIGHT
    meas[0, 1], sp[0, 1], thrust[0, 1] = [0.1, 0.2, 0.3], [0.1, 0.0, -0.1], EnvDroneParams.WEIGHT
    rpm = np.stack([
        ctbr.compute(torch.from_numpy(meas[t]), torch.from_numpy(sp[t]), torch.from_numpy(thrust[t])).numpy().copy()
        for t in range(T)
    ])
    cfg = load_yaml_config("/root/reference/configs/controllers/tracking/tracking_controller_params.yaml")
    trk = DroneTrackingController(EnvDroneParams.MASS, cfg, dt=0.04, num_envs=N, device="cpu")
    X = rng.uniform(-1, 1, (T, N, 3)).astype(np.float32) + np.array([0, 0, 1.5], np.float32)
    Q = rng.standard_normal((T, N, 4)).astype(np.float32) * 0.2 + np.array([1, 0, 0, 0], np.float32)
    Q /= np.linalg.norm(Q, axis=-1, keepdims=True)
    Xs = rng.uniform(-1, 1, (N, 3)).astype(np.float32) + np.array([0, 0, 1.5], np.float32)
    yaw = rng.uniform(-3, 3, N).astype(np.float32)
    Qs = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], 1).astype(np.float32)
    # SURVEY 8(c) tracking case in the first two rows
    X[0, 0], X[0, 1] = [0, 0, 1], [0.1, -0.2, 1.3]
    Q[0, 0] = [1, 0, 0, 0]
    Q[0, 1] = np.array([0.9, 0.1, -0.2, 0.3], np.float32) / np.linalg.norm([0.9, 0.1, -0.2, 0.3])
    X[1, :2] = X[0, :2]
    Q[1, :2] = Q[0, :2]
    Xs[0], Xs[1] = [0, 0, 1.5], [1, 1, 2.5]
    Qs[0], Qs[1] = [1, 0, 0, 0], [0.7071068, 0, 0, 0.7071068]
    out = np.stack([
        trk.compute(
            state_measurement=TrackingCtrlDroneState(X=torch.from_numpy(X[t]), Q=torch.from_numpy(Q[t])),
            state_setpoint=TrackingCtrlDroneState(X=torch.from_numpy(Xs), Q=torch.from_numpy(Qs)),
        ).numpy().copy()
        for t in range(T)
    ])
    np.savez(
        os.path.join(OUT, "controllers.npz"), mixer=ctbr.mixer.numpy(), A_inv=ctbr.A_inv.numpy(),
        inv_sqrt_kf=np.float32(ctbr.inv_sqrt_KF.item()), rate_meas=meas, rate_sp=sp, thrust_sp=thrust, rpm=rpm,
        X=X, Q=Q, X_sp=Xs, Q_sp=Qs, ctbr_cmd=out,
    )  # fmt: skip
    print("controllers.npz: hover rpm", rpm[0, 0], "tracking", out[0, :2], out[1, :2])


def make_env(action_type: int, name: str, N=32, T=200):
    import torch
    from data_driven_quad_control.envs.hover_env import HoverEnv
    from data_driven_quad_control.envs.hover_env_config import EnvActionType, get_cfgs

    cfgs = get_cfgs()
    cfgs[0]["min_hover_time_s"] = 0.2  # 5 steps: makes hover events reachable in a short fixture
    ref = HoverEnv(N, *copy.deepcopy(cfgs), device="cpu", action_type=EnvActionType(action_type), auto_target_updates=True)
    A = ref.num_actions
    acts = np.random.RandomState(action_type).uniform(-1.2, 1.2, (T, N, A)).astype(np.float32)
    acts[:, :8] = 0.0  # hovering envs: parked on their target below to exercise the hover counter
    torch.manual_seed(1234)
    ref.reset()
    rec = {k: [] for k in ("obs", "rew", "reset", "time_outs", "commands", "hover", "base_pos", "ep_len", "episode")}
    for t in range(T):
        if t % 25 == 0:
            ref.update_target_pos(torch.arange(8), ref.base_pos[:8].clone())
        o, r, d, ex = ref.step(torch.from_numpy(acts[t]))
        rec["obs"].append(o.numpy().copy())
        rec["rew"].append(r.numpy().copy())
        rec["reset"].append(d.numpy().copy())
        rec["time_outs"].append(ex["time_outs"].numpy().copy())
        rec["commands"].append(ref.commands.numpy().copy())
        rec["hover"].append(ref.hover_counter.numpy().copy())
        rec["base_pos"].append(ref.base_pos.numpy().copy())
        rec["ep_len"].append(ref.episode_length_buf.numpy().copy())
        rec["episode"].append(np.array([ex["episode"]["rew_" + k] for k in ref.reward_scales]))
    extra = {}
    if ref.uses_ctbr_actions:
        extra = dict(mixer=ref.ctbr_controller.mixer.numpy(), inv_sqrt_kf=np.float32(ref.ctbr_controller.inv_sqrt_KF.item()))
    np.savez_compressed(
        os.path.join(OUT, f"env_{name}.npz"), actions=acts, seed=1234, min_hover_time_s=0.2,
        **{k: np.stack(v) for k, v in rec.items()}, **extra,
    )  # fmt: skip
    print(f"env_{name}.npz: resets {int(np.stack(rec['reset']).sum())}, hover max {int(np.stack(rec['hover']).max())}")


def make_ddmpc():
    import ddmpc_oracle as mo

    # default sizes on PID-stabilised drone data (BASELINE configs[3] shape), 4 closed-loop solves
    u, y, st = mo.collect_initial_data(seed=0)
    prm = mo.default_params()
    y_r = np.array([1.0, 1.0, 2.5])
    ctrl = mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u, y, y_r), solve_on_init=False)
    rec = {k: [] for k in ("u_opt", "us", "ys", "y_next", "kkt")}
    for t in range(4):
        ctrl.update_and_solve_data_driven_mpc()
        k = mo.kkt_residuals(*ctrl.last["qp"], ctrl.last["z"], *ctrl.last["mult"])
        u0 = ctrl.get_optimal_control_input_at_step(0)
        st["act"](u0[None, :])
        yk = st["measure"]()[0]
        rec["u_opt"].append(ctrl.optimal_u.copy())
        rec["us"].append(ctrl.us_prev.copy())
        rec["ys"].append(ctrl.ys_prev.copy())
        rec["y_next"].append(yk)
        rec["kkt"].append([k["stationarity"], k["equality"], k["inequality"], k["min_active_multiplier"], k["n_active"]])
        ctrl.store_input_output_measurement(u0, yk)
        print("ddmpc default step", t, "u0", u0, k)
    np.savez_compressed(os.path.join(OUT, "ddmpc_default.npz"), u=u, y=y, y_r=y_r, **{k: np.stack(v) for k, v in rec.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["controllers", "env", "ddmpc"]
    if "controllers" in which:
        make_controllers()
    if "env" in which:
        make_env(2, "ctbr_fixed_yaw")
        make_env(1, "ctbr")
        make_env(0, "rotor_rpms")
    if "ddmpc" in which:
        make_ddmpc()
