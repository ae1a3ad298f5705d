"""Vectorized drone hover environment on one fused CUDA kernel.

Host mirror of the reference's `envs/hover_env.py` (`HoverEnv.__init__` :69-326, `step`
:376-547, `reset` :371-374, `reset_idx` :556-595, `get_observations` :629-630,
`get_pos/get_quat` :632-659, `update_target_pos` :667-677, `get_current_state` /
`restore_from_state` :680-753).  Constructor keywords, attributes and return values are the
reference's; every `step()` is one launch of `ddqc_env_step` (include/ddqc.h) with no host
synchronisation.

Layout: the per-env state lives in struct-of-arrays tensors of shape (k, num_envs); the
reference's (num_envs, k) attributes (`base_pos`, `base_quat`, `commands`, `last_actions`, ...)
are transposed *views* of that storage, so they alias the live state exactly like the
reference's buffers do (in-place writes through them are seen by the next step).

Documented deviations (all forced by removing the per-step host round trips):
  * `extras["episode"]["rew_*"]` are 0-dim CUDA tensors (views of a 1024-row ring written by
    the kernel) instead of Python floats; `float(x)` gives the reference's value.
  * target resampling uses a counter-based Philox stream keyed by (seed, env, event) instead
    of the global torch generator; the draw order of the reference (three batched
    `torch.rand` calls over a data-dependent index list, :328-337) cannot be reproduced
    without a device-wide scan per step.
  * derived buffers (`base_lin_vel`, `base_ang_vel`, `rel_pos`, `last_rel_pos`, `base_euler`,
    `crash_condition`, `last_base_pos`) are written every step only when
    `materialize_buffers` is on (default: num_envs <= 65536); otherwise the body-frame
    velocities and `rel_pos` are computed on access and the others are unavailable.
  * rendering (`show_viewer`, camera, debug spheres) is out of scope: `show_viewer=True` or
    `env_cfg["visualize_camera"]` raises.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Any

import numpy as np
import torch

from data_driven_quad_control_b200 import _lib
from data_driven_quad_control_b200.controllers.ctbr.ctbr_controller import build_ctbr_matrices
from data_driven_quad_control_b200.utilities.vectorized_pid_controller import VectorizedControllerState

from .hover_env_config import (
    EnvActionBounds,
    EnvActionType,
    EnvCTBRControllerConfig,
    EnvDroneParams,
    EnvState,
)

REWARD_NAMES = ("target", "closeness", "hover_time", "smooth", "yaw", "angular", "crash")

# cf2x.urdf prop{0..3}_link origins (x, y); the restated rigid-body model applies each rotor's
# thrust at these points (see DESIGN.md "physics restatement").
_ROTOR_XY = ((0.028, -0.028), (-0.028, -0.028), (-0.028, 0.028), (0.028, 0.028))
_GRAVITY = 9.81


def _f32(x: float) -> float:
    return float(np.float32(x))


def make_env_params(
    env_cfg: dict[str, Any],
    obs_cfg: dict[str, Any],
    reward_scales_dt: dict[str, float],
    yaw_lambda: float,
    command_cfg: dict[str, Any],
    action_type: EnvActionType,
    auto_target_updates: bool,
    need_euler: bool,
    seed: int,
    mixer: torch.Tensor | None = None,
    inv_sqrt_kf: float = 0.0,
    rate_pid_gains: list[list[float]] | None = None,
    independent_envs: bool = False,
) -> _lib.EnvParams:
    """Fill the C parameter block.  Pure host code (no CUDA needed); every float is the fp32
    rounding of the double the reference computes in Python."""
    P = _lib.EnvParams()
    dt = env_cfg["dt"]
    decimation = env_cfg["decimation"]
    step_dt = dt * decimation
    P.action_type = action_type.value
    P.decimation = int(decimation)
    P.max_episode_length = math.ceil(env_cfg["episode_length_s"] / step_dt)
    P.min_hover_steps = math.ceil(env_cfg["min_hover_time_s"] / step_dt)
    flags = 0
    if env_cfg["simulate_action_latency"]:
        flags |= _lib.F_ACTION_LATENCY
    if auto_target_updates:
        flags |= _lib.F_AUTO_TARGET
    if need_euler:
        flags |= _lib.F_NEED_EULER
    if independent_envs:
        flags |= _lib.F_INDEPENDENT_ENVS
    if env_cfg["termination_if_close_to_ground"] <= 0.0:
        # the reference's floor (hover_env.py:199) is not modelled: count what it would have stopped (see ddqc.h)
        flags |= _lib.F_GROUND_WATCH
    P.flags = flags
    P.clip_actions = env_cfg["clip_actions"]

    A = EnvActionType.get_num_actions(action_type)
    if action_type == EnvActionType.ROTOR_RPMS:
        lo = [EnvActionBounds.MIN_RPM] * 4
        hi = [EnvActionBounds.MAX_RPM] * 4
    else:
        lo = [0.0] + [-w for w in EnvActionBounds.MAX_ANG_VELS[: A - 1]]
        hi = [EnvActionBounds.MAX_THRUST] + list(EnvActionBounds.MAX_ANG_VELS[: A - 1])
    for j in range(A):
        P.act_lo[j] = lo[j]
        # fp32(hi) - fp32(lo) evaluated in fp32, like `(y_max - y_min)` on the bounds tensors
        P.act_span[j] = float(np.float32(hi[j]) - np.float32(lo[j]))

    P.step_dt = step_dt
    P.inv_step_dt = float(np.float32(1.0) / np.float32(step_dt))
    if action_type != EnvActionType.ROTOR_RPMS:
        assert mixer is not None and rate_pid_gains is not None
        for j in range(3):
            P.pid_kp[j], P.pid_ki[j], P.pid_kd[j] = rate_pid_gains[j]
        flat = mixer.detach().to("cpu", torch.float32).flatten().tolist()
        for j in range(16):
            P.mixer[j] = flat[j]
        P.inv_sqrt_kf = inv_sqrt_kf

    # rigid-body constants: the double-precision expressions of oracle/drone_physics.PhysParams
    inertia = EnvDroneParams.INERTIA
    jx, jy, jz = inertia["Jxx"], inertia["Jyy"], inertia["Jzz"]
    mass = EnvDroneParams.MASS
    P.dt, P.half_dt, P.dt_g = dt, dt / 2.0, dt * _GRAVITY
    P.kf = EnvDroneParams.KF
    P.dt_over_m, P.two_dt_over_m = dt / mass, 2.0 * dt / mass
    P.k_w[0], P.k_w[1], P.k_w[2] = dt / jx, dt / jy, dt / jz
    P.b_w[0], P.b_w[1], P.b_w[2] = dt * (jz - jy) / jx, dt * (jx - jz) / jy, dt * (jy - jx) / jz
    for r in range(4):
        x, y = _ROTOR_XY[r]
        P.arm_x[r], P.arm_y[r] = y, -x
        P.km_spin[r] = EnvDroneParams.KM * float(EnvDroneParams.ROTOR_SPIN_DIRECTIONS[r])
    for k in range(5):
        P.cos_coef[k] = (-1.0) ** k / math.factorial(2 * k)
        P.sinc_coef[k] = (-1.0) ** k / math.factorial(2 * k + 1)

    P.actuator_noise_std = env_cfg["actuator_noise_std"]
    P.min_rpm, P.max_rpm = EnvActionBounds.MIN_RPM, EnvActionBounds.MAX_RPM
    P.obs_noise_std = obs_cfg["obs_noise_std"]
    P.term_pitch = env_cfg["termination_if_pitch_greater_than"]
    P.term_roll = env_cfg["termination_if_roll_greater_than"]
    P.term_x = env_cfg["termination_if_x_greater_than"]
    P.term_y = env_cfg["termination_if_y_greater_than"]
    P.term_z = env_cfg["termination_if_z_greater_than"]
    P.term_ground = env_cfg["termination_if_close_to_ground"]
    P.term_ang_vel = env_cfg["termination_if_ang_vel_greater_than"]
    P.term_lin_vel = env_cfg["termination_if_lin_vel_greater_than"]
    iq = env_cfg["base_init_quat"]
    for j in range(3):
        P.init_pos[j] = env_cfg["base_init_pos"][j]
    for j in range(4):
        P.init_quat[j] = iq[j]
        P.inv_init_quat[j] = iq[j] if j == 0 else -iq[j]
    P.at_target_threshold = env_cfg["at_target_threshold"]
    P.at_target_threshold_sq = env_cfg["at_target_threshold"] ** 2
    P.inv_at_target_threshold_sq = float(np.float32(1.0) / np.float32(env_cfg["at_target_threshold"] ** 2))
    P.inv_pi_lit = float(np.float32(1.0) / np.float32(3.14159))
    P.inv_180 = float(np.float32(1.0) / np.float32(180.0))
    for j, key in enumerate(("pos_x_range", "pos_y_range", "pos_z_range")):
        lo_c, hi_c = command_cfg[key]
        P.cmd_lo[j] = lo_c
        P.cmd_span[j] = hi_c - lo_c
    scales = obs_cfg["obs_scales"]
    P.obs_scale_pos, P.obs_scale_lin, P.obs_scale_ang = scales["rel_pos"], scales["lin_vel"], scales["ang_vel"]
    for k, name in enumerate(REWARD_NAMES):
        P.rew_scale[k] = reward_scales_dt.get(name, 0.0)
    P.yaw_lambda = yaw_lambda
    P.episode_length_s = env_cfg["episode_length_s"]
    P.seed = seed & 0xFFFFFFFFFFFFFFFF
    return P


class _FusedCTBRController:
    """`env.ctbr_controller`: the rate PID + mixer run inside the fused step kernel; this object
    exposes the reference controller's matrices and its state API (ctbr_controller.py:266-292)
    on top of the env's SoA PID buffers."""

    def __init__(self, env: "HoverEnv", mats: dict[str, torch.Tensor], gains: list[list[float]]):
        self._env = env
        self.num_envs, self.device, self.dt = env.num_envs, env.device, env.step_dt
        for name, t in mats.items():
            setattr(self, name, t.to(env.device))
        self.rate_pid_params = torch.as_tensor(gains, dtype=torch.float, device=env.device)

    def reset(self) -> None:
        self._env._pid_integral.zero_()
        self._env._pid_prev_error.zero_()

    def get_state(self) -> VectorizedControllerState:
        self._env._flush_pid_reset()
        return VectorizedControllerState(
            integral=self._env._pid_integral.T.clone(), prev_error=self._env._pid_prev_error.T.clone()
        )

    def load_state(self, state: VectorizedControllerState) -> None:
        env = self._env
        env._flush_pid_reset()
        env._pid_integral.copy_(state.integral.to(env.device, torch.float32).T)
        env._pid_prev_error.copy_(state.prev_error.to(env.device, torch.float32).T)


class HoverEnv:
    def __init__(
        self,
        num_envs: int,
        env_cfg: dict[str, Any],
        obs_cfg: dict[str, Any],
        reward_cfg: dict[str, Any],
        command_cfg: dict[str, Any],
        show_viewer: bool = False,
        device: torch.device | str = "cuda",
        auto_target_updates: bool = True,
        action_type: EnvActionType = EnvActionType.CTBR,
        drone_colors: list[tuple[float, float, float, float]] | None = None,
        camera_config: dict[str, Any] | None = None,
        *,
        seed: int = 0,
        materialize_buffers: bool | None = None,
        independent_envs: bool = False,
    ):
        if show_viewer or env_cfg.get("visualize_camera", False):
            raise NotImplementedError("rendering (viewer / camera) is outside the B200 hot path")
        if drone_colors and len(drone_colors) != num_envs:
            raise ValueError("The number of colors in `drone_colors` must be equal to the number of environments.")
        self._lib = _lib.load()
        self._is_closed = False
        self.device = _lib.require_cuda(device)

        self.dt = env_cfg["dt"]
        self.decimation = env_cfg["decimation"]
        self.step_dt = self.dt * self.decimation
        self.num_envs = N = int(num_envs)
        self.action_type = action_type
        self.num_obs = EnvActionType.get_num_obs(action_type)
        self.num_actions = A = EnvActionType.get_num_actions(action_type)
        self.uses_ctbr_actions = action_type in (EnvActionType.CTBR, EnvActionType.CTBR_FIXED_YAW)
        self.num_commands = command_cfg["num_commands"]
        self.simulate_action_latency = env_cfg["simulate_action_latency"]
        self.max_episode_length = math.ceil(env_cfg["episode_length_s"] / self.step_dt)
        self.env_cfg, self.obs_cfg, self.reward_cfg, self.command_cfg = env_cfg, obs_cfg, reward_cfg, command_cfg
        self.obs_scales = obs_cfg["obs_scales"]
        self.reward_scales = reward_cfg["reward_scales"]
        self.actuator_noise_std: float = env_cfg["actuator_noise_std"]
        self.obs_noise_std: float = obs_cfg["obs_noise_std"]
        self.min_hover_time_s = env_cfg["min_hover_time_s"]
        self.min_hover_steps = math.ceil(self.min_hover_time_s / self.step_dt)
        self.at_target_threshold = env_cfg["at_target_threshold"]
        self.at_target_threshold_square = self.at_target_threshold**2
        self.auto_target_updates = auto_target_updates
        self.drone_colors_enabled = False
        self.render = False
        self.target = None

        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)

        # reward scales are multiplied by step_dt IN THE CALLER'S DICT (hover_env.py:246-247)
        for name in self.reward_scales.keys():
            if name not in REWARD_NAMES:
                raise ValueError(f"unknown reward term {name!r}")
            self.reward_scales[name] *= self.step_dt
        if list(self.reward_scales.keys()) != [n for n in REWARD_NAMES if n in self.reward_scales]:
            raise ValueError("reward_scales must list its terms in the reference order " + str(REWARD_NAMES))

        # action bounds (hover_env.py:119-151)
        if self.uses_ctbr_actions:
            max_w = EnvActionBounds.MAX_ANG_VELS[: A - 1]
            bounds = [[0.0, EnvActionBounds.MAX_THRUST]] + [[-w, w] for w in max_w]
        else:
            bounds = [[EnvActionBounds.MIN_RPM, EnvActionBounds.MAX_RPM]]
        self.action_bounds = torch.tensor(bounds, **f32)

        mats = gains = None
        if self.uses_ctbr_actions:
            mats = build_ctbr_matrices(EnvDroneParams.get())
            gains = EnvCTBRControllerConfig.get()["ctbr_controller_params"]["rate_pid_gains"]

        self._aux = (N <= 65536) if materialize_buffers is None else bool(materialize_buffers)
        yaw_scale = self.reward_scales.get("yaw", 0.0)
        need_euler = (
            self._aux
            or yaw_scale != 0.0
            # |roll| <= fp32(pi)*fp32(180/pi), which rounds to exactly 180.0f, and |pitch| <= 90:
            # with thresholds >= 180 the two predicates (hover_env.py:483-488) can never fire
            or env_cfg["termination_if_roll_greater_than"] < 180
            or env_cfg["termination_if_pitch_greater_than"] < 180
        )
        self._params = make_env_params(
            env_cfg, obs_cfg, self.reward_scales, reward_cfg["yaw_lambda"], command_cfg, action_type,
            auto_target_updates, need_euler, seed,
            mixer=mats["mixer"] if mats else None,
            inv_sqrt_kf=float(mats["inv_sqrt_KF"]) if mats else 0.0,
            rate_pid_gains=gains,
            independent_envs=independent_envs,
        )  # fmt: skip
        # independent_envs: env b behaves like its own `HoverEnv(num_envs=1)` (a reset zeroes the rate PID and refreshes
        # rel_pos of that env only) - an evaluation run of the reference's grid search as it is with num_processes = 1
        self.independent_envs = bool(independent_envs)

        # ---- persistent SoA state -------------------------------------------------------------
        self._pos = torch.zeros((3, N), **f32)
        self._quat = torch.zeros((4, N), **f32)
        self._quat[0] = 1.0
        self._vel = torch.zeros((3, N), **f32)
        self._ang = torch.zeros((3, N), **f32)
        self._commands = torch.zeros((3, N), **f32)
        self._last_actions = torch.zeros((A, N), **f32)
        self._pid_integral = torch.zeros((3, N), **f32)
        self._pid_prev_error = torch.zeros((3, N), **f32)
        self._episode_sums = torch.zeros((_lib.NUM_REWARDS, N), **f32)
        self._episode_length = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.hover_counter = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.episode_sums = {name: self._episode_sums[REWARD_NAMES.index(name)] for name in self.reward_scales}

        # ---- outputs ----------------------------------------------------------------------------
        # obs | rew | time_outs | reset live in ONE allocation (`output_block`, uint8), so that a caller who wants a
        # step's results on the host fetches them with a single device-to-host copy instead of three
        o_rew = N * self.num_obs * 4
        o_tmo = o_rew + N * 4
        o_rst = o_tmo + N * 4
        self.output_block = torch.zeros((o_rst + N,), dtype=torch.uint8, device=dev)
        self.obs_buf = self.output_block[:o_rew].view(torch.float32).view(N, self.num_obs)
        self.rew_buf = self.output_block[o_rew:o_tmo].view(torch.float32)
        self._time_outs = self.output_block[o_tmo:o_rst].view(torch.float32)
        self.reset_buf = self.output_block[o_rst:].view(torch.bool)
        self.reset_buf.fill_(True)
        self.output_layout = {"obs": (0, (N, self.num_obs), "float32"), "rew": (o_rew, (N,), "float32"),
                              "time_outs": (o_tmo, (N,), "float32"), "reset": (o_rst, (N,), "bool")}  # fmt: skip
        self._fixup_idx = torch.zeros((N,), dtype=torch.int32, device=dev)
        self._fixup_val = torch.zeros((N, 8), **f32)
        self._blv = self._bav = None  # body-frame velocity buffers (aux or on-demand cache)
        self._rel_pos = self._last_rel_pos = self._base_euler = self._crash = self._rpm = self._last_base_pos = None
        if self._aux:
            self._blv = torch.zeros((3, N), **f32)
            self._bav = torch.zeros((3, N), **f32)
            self._rel_pos = torch.zeros((3, N), **f32)
            self._last_rel_pos = torch.zeros((3, N), **f32)
            self._base_euler = torch.zeros((3, N), **f32)
            self._crash = torch.zeros((N,), dtype=torch.bool, device=dev)
            self._rpm = torch.zeros((4, N), **f32)
            self._last_base_pos = torch.zeros((3, N), **f32)
        self._body_valid = self._aux  # are _blv/_bav current?
        self._meas_override_next = False

        # ---- control block ----------------------------------------------------------------------
        self._ctl = torch.zeros((C.sizeof(_lib.EnvCtl),), dtype=torch.uint8, device=dev)
        ring_off = _lib.EnvCtl.ep_ring.offset
        self._ep_ring = self._ctl[ring_off : ring_off + _lib.EP_RING * 8 * 4].view(torch.float32).view(_lib.EP_RING, 8)
        self._ep_row = 0
        self._ep_dicts: dict[int, dict[str, torch.Tensor]] = {}

        self._state = _lib.EnvState(
            pos=self._pos.data_ptr(), quat=self._quat.data_ptr(), vel=self._vel.data_ptr(), ang=self._ang.data_ptr(),
            commands=self._commands.data_ptr(), last_actions=self._last_actions.data_ptr(),
            pid_integral=self._pid_integral.data_ptr(), pid_prev_error=self._pid_prev_error.data_ptr(),
            episode_sums=self._episode_sums.data_ptr(), episode_length=self._episode_length.data_ptr(),
            hover_counter=self.hover_counter.data_ptr(),
        )  # fmt: skip
        self._out = self._make_out()

        self.base_init_pos = torch.tensor(env_cfg["base_init_pos"], **f32)
        self.base_init_quat = torch.tensor(env_cfg["base_init_quat"], **f32)
        self.inv_base_init_quat = self.base_init_quat * torch.tensor([1.0, -1.0, -1.0, -1.0], **f32)
        self.body_rate_setpoints = torch.zeros((N, 3), **f32)  # kept for API parity; the kernel recomputes it

        self.extras: dict[str, Any] = {"observations": {}}
        if self.uses_ctbr_actions:
            self.ctbr_controller = _FusedCTBRController(self, mats, gains)
        self._idx_all = torch.arange(N, device=dev)

    # ------------------------------------------------------------------------------ plumbing
    def _make_out(self, force_body: bool = False) -> _lib.EnvOut:
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        body = self._aux or force_body
        return _lib.EnvOut(
            obs=self.obs_buf.data_ptr(), rew=self.rew_buf.data_ptr(), reset=self.reset_buf.data_ptr(),
            time_outs=self._time_outs.data_ptr(),
            base_lin_vel=ptr(self._blv) if body else None, base_ang_vel=ptr(self._bav) if body else None,
            rel_pos=ptr(self._rel_pos), last_rel_pos=ptr(self._last_rel_pos), base_euler=ptr(self._base_euler),
            crash=ptr(self._crash), rpm=ptr(self._rpm), last_base_pos=ptr(self._last_base_pos),
            fixup_idx=self._fixup_idx.data_ptr(), fixup_val=self._fixup_val.data_ptr(),
        )  # fmt: skip

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _next_episode_row(self) -> int:
        """Row of the extras["episode"] ring the next launch writes (passed down as DdqcEnvOut.ep_row: the host, not
        a device counter, names the row, so the views handed out here stay right under CUDA-graph replay)."""
        row = (self._ep_row + 1) % _lib.EP_RING
        self._out.ep_row = row
        return row

    def _advance_episode_row(self) -> None:
        self._ep_row = (self._ep_row + 1) % _lib.EP_RING
        d = self._ep_dicts.get(self._ep_row)
        if d is None:
            row = self._ep_ring[self._ep_row]
            d = {"rew_" + name: row[REWARD_NAMES.index(name)] for name in self.reward_scales}
            self._ep_dicts[self._ep_row] = d
        self.extras["episode"] = d

    def _flush_pid_reset(self) -> None:
        if self.uses_ctbr_actions:
            with _lib.device_guard(self.device):
                rc = self._lib.ddqc_env_flush_pid_reset(C.byref(self._state), self._ctl.data_ptr(), self.num_envs, self._stream())
            _lib.check(rc, "ddqc_env_flush_pid_reset")

    def _ensure_body_vel(self) -> None:
        if self._body_valid:
            return
        if self._blv is None:
            self._blv = torch.empty((3, self.num_envs), dtype=torch.float32, device=self.device)
            self._bav = torch.empty((3, self.num_envs), dtype=torch.float32, device=self.device)
        with _lib.device_guard(self.device):
            rc = self._lib.ddqc_env_body_vel(
                C.byref(self._state), self.num_envs, self._blv.data_ptr(), self._bav.data_ptr(), self._stream()
            )
        _lib.check(rc, "ddqc_env_body_vel")
        self._body_valid = True

    def _as_idx(self, envs_idx) -> torch.Tensor:
        idx = torch.as_tensor(envs_idx, device=self.device)
        if idx.dtype == torch.bool:
            idx = idx.nonzero(as_tuple=False).flatten()
        return idx.to(torch.int64).reshape(-1).contiguous()

    # ------------------------------------------------------------------------------ reference attributes
    @property
    def base_pos(self) -> torch.Tensor:
        return self._pos.T

    @property
    def base_quat(self) -> torch.Tensor:
        return self._quat.T

    @property
    def commands(self) -> torch.Tensor:
        return self._commands.T

    @property
    def last_actions(self) -> torch.Tensor:
        return self._last_actions.T

    @property
    def actions(self) -> torch.Tensor:
        # after step(): last_actions[:] = actions (hover_env.py:544), so the two coincide
        return self._last_actions.T

    @property
    def base_lin_vel(self) -> torch.Tensor:
        self._ensure_body_vel()
        return self._blv.T

    @property
    def base_ang_vel(self) -> torch.Tensor:
        self._ensure_body_vel()
        return self._bav.T

    @property
    def rel_pos(self) -> torch.Tensor:
        if self._rel_pos is not None:
            return self._rel_pos.T
        return (self._commands - self._pos).T

    def _aux_attr(self, t, name):
        if t is None:
            raise AttributeError(f"`{name}` is only kept when the env is built with materialize_buffers=True")
        return t

    @property
    def last_rel_pos(self) -> torch.Tensor:
        return self._aux_attr(self._last_rel_pos, "last_rel_pos").T

    @property
    def last_base_pos(self) -> torch.Tensor:
        return self._aux_attr(self._last_base_pos, "last_base_pos").T

    @property
    def base_euler(self) -> torch.Tensor:
        return self._aux_attr(self._base_euler, "base_euler").T

    @property
    def crash_condition(self) -> torch.Tensor:
        return self._aux_attr(self._crash, "crash_condition")

    @property
    def rotor_rpms(self) -> torch.Tensor:
        return self._aux_attr(self._rpm, "rotor_rpms").T

    @property
    def episode_length_buf(self) -> torch.Tensor:
        return self._episode_length

    @episode_length_buf.setter
    def episode_length_buf(self, value: torch.Tensor) -> None:
        # rsl_rl assigns a fresh randint tensor (init_at_random_ep_len); keep our storage
        self._episode_length.copy_(torch.as_tensor(value, device=self.device).to(torch.int32))

    # ------------------------------------------------------------------------------ API
    def reset(self) -> tuple[torch.Tensor, None]:
        self.reset_buf[:] = True
        self.reset_idx(self._idx_all)
        return self.obs_buf, None

    def step(self, actions: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, dict[str, Any]]:
        a = actions
        if a.dtype != torch.float32 or a.device != self.device or not a.is_contiguous():
            a = a.to(device=self.device, dtype=torch.float32).contiguous()
        if a.shape != (self.num_envs, self.num_actions):
            raise ValueError(f"actions must have shape {(self.num_envs, self.num_actions)}")
        override = None
        if self._meas_override_next:
            self._ensure_body_vel()
            override = self._bav.data_ptr()
            self._meas_override_next = False
        self._next_episode_row()
        with _lib.device_guard(self.device):
            rc = self._lib.ddqc_env_step(
                C.byref(self._params), C.byref(self._state), C.byref(self._out), self._ctl.data_ptr(),
                a.data_ptr(), override, self.num_envs, self._stream(),
            )  # fmt: skip
        _lib.check(rc, "ddqc_env_step")
        self._body_valid = self._aux
        self.extras["time_outs"] = self._time_outs
        self._advance_episode_row()
        return self.obs_buf, self.rew_buf, self.reset_buf, self.extras

    def reset_idx(self, envs_idx: torch.Tensor) -> None:
        idx = self._as_idx(envs_idx)
        if idx.numel() == 0:
            return
        cmd_old = self._commands.clone() if self._aux else None
        self._next_episode_row()
        with _lib.device_guard(self.device):
            rc = self._lib.ddqc_env_reset_idx(
                C.byref(self._params), C.byref(self._state), C.byref(self._out), self._ctl.data_ptr(),
                idx.data_ptr(), idx.numel(), self.num_envs, self._stream(),
            )  # fmt: skip
        _lib.check(rc, "ddqc_env_reset_idx")
        self._advance_episode_row()
        if self._aux:
            # reference buffers touched by reset_idx (hover_env.py:561-574): all-env rel_pos refresh
            # against the commands held BEFORE the resample at the end of reset_idx
            self._last_base_pos[:, idx] = self.base_init_pos[:, None]
            self._rel_pos.copy_(cmd_old - self._pos)
            self._last_rel_pos.copy_(cmd_old - self._last_base_pos)
            self._blv[:, idx] = 0.0
            self._bav[:, idx] = 0.0
        elif self._blv is not None and self._body_valid:
            self._blv[:, idx] = 0.0
            self._bav[:, idx] = 0.0

    def get_observations(self) -> tuple[torch.Tensor, dict[str, Any]]:
        return self.obs_buf, self.extras

    def compute_observations(self) -> torch.Tensor:
        """The observation of the last step (hover_env.py:597-627); it is produced by the step kernel."""
        return self.obs_buf

    def get_pos(self, add_noise: bool = True) -> torch.Tensor:
        if add_noise and self.obs_noise_std > 0.0:
            noise = torch.randn((self.num_envs, 3), dtype=torch.float32, device=self.device)
            return self.base_pos + noise * self.obs_noise_std / self.obs_scales["rel_pos"]
        return self.base_pos

    def get_quat(self, add_noise: bool = True) -> torch.Tensor:
        if add_noise and self.obs_noise_std > 0.0:
            quat = self._add_noise(self.base_quat, self.obs_noise_std)
            return quat / quat.norm(dim=1, keepdim=True)
        return self.base_quat

    def _add_noise(self, input_tensor: torch.Tensor, noise_std: float) -> torch.Tensor:
        return input_tensor + torch.randn(input_tensor.shape, dtype=input_tensor.dtype, device=self.device) * noise_std

    def update_target_pos(self, envs_idx: torch.Tensor, target_pos: torch.Tensor) -> None:
        idx = self._as_idx(envs_idx)
        tgt = torch.as_tensor(target_pos, device=self.device).to(torch.float32).contiguous()
        broadcast = 1 if tgt.numel() == 3 else 0
        if not broadcast and tgt.shape != (idx.numel(), 3):
            raise ValueError("target_pos must have shape (len(envs_idx), 3) or (3,)")
        with _lib.device_guard(self.device):
            rc = self._lib.ddqc_env_update_target(
                C.byref(self._state), idx.data_ptr(), idx.numel(), self.num_envs, tgt.data_ptr(), broadcast, self._stream()
            )
        _lib.check(rc, "ddqc_env_update_target")

    def get_current_state(self, envs_idx: torch.Tensor) -> EnvState:
        idx = self._as_idx(envs_idx)
        k, A, dev = idx.numel(), self.num_actions, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        if k == 0:  # an empty selection is an empty state (nothing to launch)
            return EnvState(
                base_pos=torch.empty((0, 3), **f32), base_quat=torch.empty((0, 4), **f32),
                base_lin_vel=torch.empty((0, 3), **f32), base_ang_vel=torch.empty((0, 3), **f32),
                commands=torch.empty((0, 3), **f32), episode_length=torch.empty((0,), dtype=torch.int32, device=dev),
                last_actions=torch.empty((0, A), **f32),
                ctbr_controller_state=self.ctbr_controller.get_state() if self.uses_ctbr_actions else None,
            )  # fmt: skip
        out = EnvState(
            base_pos=torch.empty((k, 3), **f32), base_quat=torch.empty((k, 4), **f32),
            base_lin_vel=torch.empty((k, 3), **f32), base_ang_vel=torch.empty((k, 3), **f32),
            commands=torch.empty((k, 3), **f32), episode_length=torch.empty((k,), dtype=torch.int32, device=dev),
            last_actions=torch.empty((k, A), **f32),
            ctbr_controller_state=self.ctbr_controller.get_state() if self.uses_ctbr_actions else None,
        )  # fmt: skip
        use_buf = self._body_valid and self._bav is not None
        with _lib.device_guard(self.device):
            rc = self._lib.ddqc_env_get_state(
                C.byref(self._params), C.byref(self._state), idx.data_ptr(), k, self.num_envs,
                out.base_pos.data_ptr(), out.base_quat.data_ptr(), out.base_lin_vel.data_ptr(),
                out.base_ang_vel.data_ptr(), out.commands.data_ptr(), out.episode_length.data_ptr(),
                out.last_actions.data_ptr(), self._bav.data_ptr() if use_buf else None,
                self._blv.data_ptr() if use_buf else None, self._stream(),
            )  # fmt: skip
        _lib.check(rc, "ddqc_env_get_state")
        return out

    def restore_from_state(self, envs_idx: torch.Tensor, saved_state: EnvState) -> None:
        idx = self._as_idx(envs_idx)
        k, dev = idx.numel(), self.device
        cast = lambda t, cols: t.to(device=dev, dtype=torch.float32).reshape(-1, cols).contiguous()  # noqa: E731
        pos, quat = cast(saved_state.base_pos, 3), cast(saved_state.base_quat, 4)
        lin, ang = cast(saved_state.base_lin_vel, 3), cast(saved_state.base_ang_vel, 3)
        cmd, act = cast(saved_state.commands, 3), cast(saved_state.last_actions, self.num_actions)
        ep = saved_state.episode_length.to(device=dev, dtype=torch.int32).reshape(-1).contiguous()
        for t in (pos, quat, lin, ang, cmd, act):
            if t.shape[0] != k:
                raise ValueError("saved_state does not match the number of env indices")
        if k == 0:  # nothing is launched (and no ring row is consumed) for an empty index list
            return
        # body-frame buffers must be valid for every env before the restored rows are patched in
        self._ensure_body_vel()
        out = self._make_out(force_body=True)
        out.ep_row = self._next_episode_row()
        with _lib.device_guard(self.device):
            rc = self._lib.ddqc_env_restore_state(
                C.byref(self._params), C.byref(self._state), C.byref(out), self._ctl.data_ptr(), idx.data_ptr(), k,
                self.num_envs, pos.data_ptr(), quat.data_ptr(), lin.data_ptr(), ang.data_ptr(), cmd.data_ptr(),
                ep.data_ptr(), act.data_ptr(), self._stream(),
            )  # fmt: skip
        _lib.check(rc, "ddqc_env_restore_state")
        self._advance_episode_row()
        self._meas_override_next = True  # next step measures the restored body rates (hover_env.py:729-735)
        if self.uses_ctbr_actions:
            assert saved_state.ctbr_controller_state is not None
            self.ctbr_controller.load_state(saved_state.ctbr_controller_state)

    def close(self) -> None:
        self._is_closed = True

    def check_ground_plane(self, what: str = "run") -> int:
        """Number of env-steps that ended below z = 0 (synchronises).  The reference's scene has a ground plane with
        collision (hover_env.py:199) which this restated rigid body does not model; with
        `termination_if_close_to_ground <= 0` a drone can therefore sink through the floor here where the reference
        would leave it lying on it.  A non-zero count raises a `RuntimeWarning`: the results of such a run cannot be
        compared with the reference's."""
        import warnings

        k = int(self.rollout_stats()["below_ground_env_steps"])
        if k:
            warnings.warn(f"{what}: {k} env-steps ended below the ground plane z = 0, which is not modelled "
                          "(the reference's drone would be in contact with the floor); results diverge from the reference's",
                          RuntimeWarning, stacklevel=2)  # fmt: skip
        return k

    # ------------------------------------------------------------------------------ rollout statistics
    def rollout_stats(self) -> dict[str, float]:
        """Totals accumulated on the device since construction (one D2H copy, synchronises):
        env steps, resets, time-outs, sum of rewards, per-term sums of finished episodes.  These are
        the quantities all-reduced across ranks by `parallel.allreduce_rollout_stats`."""
        raw = self._ctl[: _lib.EnvCtl.blocks_done.offset].cpu().numpy().tobytes()
        ctl = _lib.EnvCtl.from_buffer_copy(raw + bytes(C.sizeof(_lib.EnvCtl) - len(raw)))
        stats = {
            "steps": float(ctl.step_count),
            "env_steps": float(ctl.total_env_steps),
            "resets": float(ctl.total_resets),
            "time_outs": float(ctl.total_timeouts),
            "reward_sum": float(ctl.total_reward),
            # env-steps that ended below the (unmodelled) ground plane; only counted when the configuration disables
            # the low-altitude termination.  Non-zero = the run left the regime in which it can agree with the reference
            "below_ground_env_steps": float(ctl.total_below_ground),
        }
        for k, name in enumerate(REWARD_NAMES):
            stats["episode_sum_" + name] = float(ctl.total_episode_sums[k])
        return stats
