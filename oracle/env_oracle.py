"""TEST INFRASTRUCTURE ONLY (oracle) -- numpy fp32 restatement of the vectorized
drone hover environment step.  Never imported by the product package.

Follows (reference `src/data_driven_quad_control/`):
  envs/hover_env.py:376-547   step            (quirks A1-A20 of SURVEY.md Appendix A)
  envs/hover_env.py:556-595   reset_idx       (A16-A18)
  envs/hover_env.py:344-369   _hovering_at_target
  envs/hover_env.py:328-342   _resample_commands
  envs/hover_env.py:597-627   compute_observations
  envs/hover_env.py:680-753   get_current_state / restore_from_state
  envs/hover_env.py:816-860   reward terms
  controllers/ctbr/ctbr_controller.py:221-264  CTBR rate PID + mixer + RPM
  utilities/vectorized_pid_controller.py:97-126  PID
  utilities/math_utils.py:13-20  linear_interpolate
and, for the part Genesis performs, `oracle/drone_physics.py` (restated; parity
unpinned there).

Pinning: `tests/test_oracle_reference_parity.py` runs the reference's own
`hover_env.py` on `oracle/genesis_shim` in the build container and checks this
oracle against it (reset indices exact, floats to a few ulp); the frozen
outputs live in `tests/golden/env_*.npz` (made by `oracle/make_golden.py`).

All arithmetic is elementwise float32 in a fixed order (no BLAS, no reductions;
the env/controller code the reference owns is evaluated without FMA exactly as
torch's elementwise kernels do, the restated Genesis part uses explicit fma()),
so the CUDA kernel -- compiled with -fmad=false -- reproduces it bit for bit;
the only tolerance-level differences are the libm calls (atan2/asin for the
Euler angles, exp for the yaw reward, log/cos for Gaussian noise).

Tensor / python-scalar divisions (`x / dt`, `x / 3.14159`, ...) follow torch's
CUDA semantics -- the reference's default device -- i.e. x * (1.0f / s)
(ATen div_true_kernel_cuda, cpu-scalar fast path); torch's CPU kernels divide.
The two differ in the last bit only.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from drone_physics import PhysParams, quat_mul, rotate_by_quat, rotor_wrench, substep

F32 = np.float32

ROTOR_RPMS, CTBR, CTBR_FIXED_YAW = 0, 1, 2
NUM_ACTIONS = {ROTOR_RPMS: 4, CTBR: 4, CTBR_FIXED_YAW: 3}
NUM_OBS = {ROTOR_RPMS: 17, CTBR: 17, CTBR_FIXED_YAW: 16}

# hover_env_config.py:172-180
BASE_RPM = 14468.429183500699
MIN_RPM = 0.2 * BASE_RPM
MAX_RPM = 1.8 * BASE_RPM
MAX_THRUST = 2 * (0.027 * 9.81)
MAX_ANG_VELS = (15.0, 15.0, 10.0)

REWARD_ORDER = ("target", "closeness", "hover_time", "smooth", "yaw", "angular", "crash")


# ----------------------------------------------------------------------------------
# counter-based RNG shared (by specification) with the CUDA kernels
# ----------------------------------------------------------------------------------
PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 on uint32 numpy arrays (counter words c0..c3, key k0,k1)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(PHILOX_M0) * c0
        p1 = np.uint64(PHILOX_M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(PHILOX_W0)) & mask
        k1 = (k1 + np.uint64(PHILOX_W1)) & mask
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def u32_to_unit(x):
    """24 high bits -> [0,1) float32 (the same mapping torch's CPU uniform uses)."""
    return (x >> np.uint32(8)).astype(F32) * F32(1.0 / 16777216.0)


# Philox stream ids (counter word 3)
STREAM_HOVER, STREAM_RESET, STREAM_ACT_NOISE, STREAM_OBS_NOISE = 0, 1, 2, 3


class PhiloxRng:
    """Per-env counter RNG: counter = (env, event_lo, event_hi, stream[+k<<8]), key = seed."""

    def __init__(self, seed: int):
        self.k0 = seed & 0xFFFFFFFF
        self.k1 = (seed >> 32) & 0xFFFFFFFF
        self.event = 0  # advanced by 1 per step() / external reset call

    def _words(self, idx, stream, sub=0):
        idx = np.asarray(idx, dtype=np.uint32)
        ev_lo = np.full(idx.shape, self.event & 0xFFFFFFFF, dtype=np.uint32)
        ev_hi = np.full(idx.shape, (self.event >> 32) & 0xFFFFFFFF, dtype=np.uint32)
        st = np.full(idx.shape, stream | (sub << 8), dtype=np.uint32)
        return philox4x32_10(idx, ev_lo, ev_hi, st, self.k0, self.k1)

    def uniform3(self, idx, stream):
        r0, r1, r2, _ = self._words(idx, stream)
        return u32_to_unit(r0), u32_to_unit(r1), u32_to_unit(r2)

    def normal(self, idx, stream, count):
        """`count` N(0,1) samples per env via Box-Muller on consecutive Philox blocks."""
        out = []
        sub = 0
        while len(out) < count:
            r = self._words(idx, stream, sub)
            sub += 1
            for a, b in ((r[0], r[1]), (r[2], r[3])):
                u1 = ((a >> np.uint32(8)).astype(F32) + F32(1.0)) * F32(1.0 / 16777216.0)  # (0,1]
                u2 = u32_to_unit(b)
                rad = np.sqrt(F32(-2.0) * np.log(u1))
                ang = F32(2.0 * math.pi) * u2
                out.append(rad * np.cos(ang))
                out.append(rad * np.sin(ang))
        return out[:count]

    def advance(self):
        self.event += 1


class TorchRng:
    """Draws exactly like the reference (`math_utils.py:4-10` via `hover_env.py:328-337`):
    three `torch.rand(len(idx))` calls on the global generator, x then y then z."""

    def __init__(self):
        import torch

        self._torch = torch

    def uniform3(self, idx, stream):
        t = self._torch
        n = len(idx)
        return tuple(t.rand(n).numpy().astype(F32) for _ in range(3))

    def normal(self, idx, stream, count):
        raise NotImplementedError("noise parity against torch.randn is statistical only")

    def advance(self):
        pass


@dataclass
class OracleEnvState:
    base_pos: np.ndarray
    base_quat: np.ndarray
    base_lin_vel: np.ndarray
    base_ang_vel: np.ndarray
    commands: np.ndarray
    episode_length: np.ndarray
    last_actions: np.ndarray
    pid_integral: np.ndarray | None
    pid_prev_error: np.ndarray | None


def _cols(a):
    return tuple(a[:, j] for j in range(a.shape[1]))


class EnvOracle:
    def __init__(
        self,
        num_envs,
        env_cfg,
        obs_cfg,
        reward_cfg,
        command_cfg,
        action_type=CTBR,
        auto_target_updates=True,
        mixer=None,
        inv_sqrt_kf=None,
        rate_pid_gains=((10.0, 0.0, 0.1),) * 3,
        rng=None,
        scale_rewards_by_dt=True,
        independent_envs=False,
    ):
        # independent_envs: every env is its own num_envs = 1 instance of the reference env, i.e. the two batch-global
        # side effects of reset_idx (hover_env.py:563-564 rel_pos of ALL envs, :592-593 PID of ALL envs) reach the
        # reset envs only - an evaluation run of the reference's grid search as it is with num_processes = 1
        self.independent_envs = bool(independent_envs)
        N = self.num_envs = int(num_envs)
        self.action_type = action_type
        self.num_actions = A = NUM_ACTIONS[action_type]
        self.num_obs = NUM_OBS[action_type]
        self.uses_ctbr = action_type in (CTBR, CTBR_FIXED_YAW)
        self.env_cfg, self.obs_cfg, self.command_cfg = env_cfg, obs_cfg, command_cfg
        self.dt = env_cfg["dt"]
        self.decimation = env_cfg["decimation"]
        self.step_dt = self.dt * self.decimation
        self.max_episode_length = math.ceil(env_cfg["episode_length_s"] / self.step_dt)
        self.min_hover_steps = math.ceil(env_cfg["min_hover_time_s"] / self.step_dt)
        self.at_target_threshold = env_cfg["at_target_threshold"]
        self.at_target_threshold_square = self.at_target_threshold**2
        self.auto_target_updates = auto_target_updates
        self.latency = env_cfg["simulate_action_latency"]
        self.actuator_noise_std = env_cfg["actuator_noise_std"]
        self.obs_noise_std = obs_cfg["obs_noise_std"]
        self.obs_scales = obs_cfg["obs_scales"]
        self.yaw_lambda = reward_cfg["yaw_lambda"]
        # hover_env.py:246-247 (scales multiplied by step_dt, in double)
        self.reward_scales = {
            k: (v * self.step_dt if scale_rewards_by_dt else v) for k, v in reward_cfg["reward_scales"].items()
        }
        self.P = PhysParams.cf2x(self.dt)
        self.rng = rng if rng is not None else PhiloxRng(0)

        # action bounds (hover_env.py:119-151), fp32
        if self.uses_ctbr:
            hi = [MAX_THRUST] + list(MAX_ANG_VELS[: A - 1])
            lo = [0.0] + [-x for x in MAX_ANG_VELS[: A - 1]]
        else:
            lo, hi = [MIN_RPM] * 4, [MAX_RPM] * 4
        self.act_lo = np.array(lo, dtype=F32)
        self.act_hi = np.array(hi, dtype=F32)
        self.act_span = self.act_hi - self.act_lo  # fp32 subtraction, as `(y_max - y_min)` on tensors

        if self.uses_ctbr:
            assert mixer is not None and inv_sqrt_kf is not None
            self.mixer = np.asarray(mixer, dtype=F32).reshape(4, 4)
            self.inv_sqrt_kf = F32(inv_sqrt_kf)
            g = np.asarray(rate_pid_gains, dtype=F32)
            self.kp, self.ki, self.kd = g[:, 0].copy(), g[:, 1].copy(), g[:, 2].copy()
            self.pid_integral = np.zeros((N, 3), F32)
            self.pid_prev_error = np.zeros((N, 3), F32)

        self.base_init_pos = np.array(env_cfg["base_init_pos"], dtype=F32)
        self.base_init_quat = np.array(env_cfg["base_init_quat"], dtype=F32)
        self.inv_base_init_quat = self.base_init_quat * np.array([1, -1, -1, -1], dtype=F32)

        z = lambda *s: np.zeros(s, F32)  # noqa: E731
        # simulator state (Genesis side)
        self.sim_pos, self.sim_quat, self.sim_vel, self.sim_ang = z(N, 3), z(N, 4), z(N, 3), z(N, 3)
        self.sim_quat[:, 0] = 1.0
        # env buffers (hover_env.py:253-311)
        self.obs_buf = z(N, self.num_obs)
        self.rew_buf = z(N)
        self.reset_buf = np.ones(N, dtype=bool)
        self.time_outs = z(N)
        self.episode_length_buf = np.zeros(N, np.int32)
        self.commands = z(N, 3)
        self.actions = z(N, A)
        self.last_actions = z(N, A)
        self.base_pos, self.base_quat = z(N, 3), z(N, 4)
        self.base_lin_vel, self.base_ang_vel = z(N, 3), z(N, 3)
        self.last_base_pos = z(N, 3)
        self.rel_pos, self.last_rel_pos = z(N, 3), z(N, 3)
        self.base_euler = z(N, 3)
        self.crash_condition = np.zeros(N, dtype=bool)
        self.hover_counter = np.zeros(N, np.int32)
        self.episode_sums = {k: z(N) for k in self.reward_scales}
        self.extras_episode = {}
        self.body_rate_setpoints = z(N, 3)
        self.last_rpm = z(N, 4)
        # env-steps that ended below the ground plane z = 0, which neither this restatement nor the product models
        # (reference: plane entity with collision, hover_env.py:199); counted only when the low-altitude termination
        # is disabled, like DDQC_F_GROUND_WATCH of the kernel
        self.below_ground_env_steps = 0

    # ------------------------------------------------------------------ helpers
    def _resample_commands(self, idx, stream):
        if len(idx) == 0:
            return
        ux, uy, uz = self.rng.uniform3(idx, stream)
        for j, (u, key) in enumerate(zip((ux, uy, uz), ("pos_x_range", "pos_y_range", "pos_z_range"))):
            lo, hi = self.command_cfg[key]
            # gs_rand_float: (upper - lower) * rand + lower
            self.commands[idx, j] = F32(hi - lo) * u + F32(lo)

    def _refresh_body_frame(self):
        q = _cols(self.base_quat)
        qc = (q[0], -q[1], -q[2], -q[3])
        lv = rotate_by_quat(*_cols(self.sim_vel), *qc)
        world_ang = rotate_by_quat(*_cols(self.sim_ang), *_cols(self.sim_quat))  # DroneEntity.get_ang()
        av = rotate_by_quat(*world_ang, *qc)
        for j in range(3):
            self.base_lin_vel[:, j] = lv[j]
            self.base_ang_vel[:, j] = av[j]

    # ------------------------------------------------------------------ reset
    def reset(self):
        self.reset_buf[:] = True
        self.reset_idx(np.arange(self.num_envs))
        return self.obs_buf

    def reset_idx(self, idx):
        """External call (hover_env.py:371-374 / direct use): one RNG event per call."""
        self._reset_idx(idx)
        self.rng.advance()

    def _reset_idx(self, idx, stream=STREAM_RESET):
        idx = np.asarray(idx, dtype=np.int64)
        if len(idx) == 0:
            return
        self.base_pos[idx] = self.base_init_pos
        self.last_base_pos[idx] = self.base_init_pos
        if self.independent_envs:
            self.rel_pos[idx] = self.commands[idx] - self.base_pos[idx]
            self.last_rel_pos[idx] = self.commands[idx] - self.last_base_pos[idx]
        else:
            self.rel_pos = self.commands - self.base_pos  # ALL envs (A16)
            self.last_rel_pos = self.commands - self.last_base_pos
        self.base_quat[idx] = self.base_init_quat
        self.sim_pos[idx] = self.base_pos[idx]
        self.sim_quat[idx] = self.base_quat[idx]
        self.sim_vel[idx] = 0
        self.sim_ang[idx] = 0
        self.base_lin_vel[idx] = 0
        self.base_ang_vel[idx] = 0
        self.last_actions[idx] = 0
        self.episode_length_buf[idx] = 0
        self.reset_buf[idx] = True
        self.hover_counter[idx] = 0
        self.extras_episode = {}
        for k in self.episode_sums:
            self.extras_episode["rew_" + k] = float(np.mean(self.episode_sums[k][idx])) / self.env_cfg["episode_length_s"]
            self.episode_sums[k][idx] = 0
        if self.uses_ctbr and self.independent_envs:
            self.pid_integral[idx] = 0
            self.pid_prev_error[idx] = 0
        elif self.uses_ctbr:  # whole-batch PID reset (A17)
            self.pid_integral[:] = 0
            self.pid_prev_error[:] = 0
        self._resample_commands(idx, stream)

    # ------------------------------------------------------------------ step
    def compute_rpm(self, exec_actions):
        N = self.num_envs
        y = self.act_lo[None, :] + ((exec_actions - F32(-1.0)) * self.act_span[None, :]) / F32(2.0)
        if not self.uses_ctbr:
            return y
        if self.action_type == CTBR:
            self.body_rate_setpoints[:, :] = y[:, 1:]
        else:
            self.body_rate_setpoints[:, :2] = y[:, 1:]
        sdt = F32(self.step_dt)
        err = self.body_rate_setpoints - self.base_ang_vel
        self.pid_integral = self.pid_integral + err * sdt
        deriv = (err - self.pid_prev_error) * (F32(1.0) / sdt)
        self.pid_prev_error = err.copy()
        acc = (self.kp[None, :] * err + self.ki[None, :] * self.pid_integral) + self.kd[None, :] * deriv
        c = (acc[:, 0], acc[:, 1], acc[:, 2], y[:, 0])
        rpm = np.zeros((N, 4), F32)
        M = self.mixer
        for i in range(4):
            f = ((M[i, 0] * c[0] + M[i, 1] * c[1]) + M[i, 2] * c[2]) + M[i, 3] * c[3]
            f = np.where(f < F32(0.0), F32(0.0), f)  # clamp(min=0), NaN stays NaN
            rpm[:, i] = np.sqrt(f) * self.inv_sqrt_kf
        return rpm

    def step(self, actions):
        N = self.num_envs
        actions = np.asarray(actions, dtype=F32).reshape(N, self.num_actions)
        c = F32(self.env_cfg["clip_actions"])
        self.actions = np.where(actions < -c, -c, np.where(actions > c, c, actions)).astype(F32)
        exec_actions = self.last_actions if self.latency else self.actions
        rpm = self.compute_rpm(exec_actions)
        if self.actuator_noise_std > 0.0:
            z = self.rng.normal(np.arange(N), STREAM_ACT_NOISE, 4)
            rpm = rpm + np.stack(z, 1) * F32(self.actuator_noise_std)
            rpm = np.clip(rpm, F32(MIN_RPM), F32(MAX_RPM))
        self.last_rpm = rpm

        p, q, v, w = _cols(self.sim_pos), _cols(self.sim_quat), _cols(self.sim_vel), _cols(self.sim_ang)
        wrench = rotor_wrench(self.P, _cols(rpm))
        for _ in range(self.decimation):
            p, q, v, w = substep(self.P, wrench, p, q, v, w)
        self.sim_pos, self.sim_quat = np.stack(p, 1), np.stack(q, 1)
        self.sim_vel, self.sim_ang = np.stack(v, 1), np.stack(w, 1)

        # buffers (hover_env.py:450-470)
        self.episode_length_buf += 1
        self.last_base_pos[:] = self.base_pos
        self.base_pos[:] = self.sim_pos
        self.rel_pos = self.commands - self.base_pos
        self.last_rel_pos = self.commands - self.last_base_pos
        self.base_quat[:] = self.sim_quat
        iq = self.inv_base_init_quat
        rw, rx, ry, rz = quat_mul(*_cols(self.base_quat), iq[0], iq[1], iq[2], iq[3])
        roll = np.arctan2((rw * rx + ry * rz) * F32(2.0), F32(1.0) - (rx * rx + ry * ry) * F32(2.0))
        pitch = np.arcsin(np.clip((rw * ry - rz * rx) * F32(2.0), F32(-1.0), F32(1.0)))
        yaw = np.arctan2((rw * rz + rx * ry) * F32(2.0), F32(1.0) - (ry * ry + rz * rz) * F32(2.0))
        self.base_euler = np.stack((roll, pitch, yaw), 1) * F32(180.0 / math.pi)
        self._refresh_body_frame()

        # automatic target updates (hover_env.py:473-478, 344-369)
        self.hover_resampled = np.zeros(N, dtype=bool)
        if self.auto_target_updates:
            rp = self.rel_pos
            dist = np.sqrt((rp[:, 0] * rp[:, 0] + rp[:, 1] * rp[:, 1]) + rp[:, 2] * rp[:, 2])
            at_target = dist < F32(self.at_target_threshold)
            if self.min_hover_steps > 0:
                self.hover_counter[at_target] += 1
                self.hover_counter[~at_target] = 0
                idx = np.nonzero(self.hover_counter >= self.min_hover_steps)[0]
            else:
                idx = np.nonzero(at_target)[0]
            if len(idx) > 0:
                self._resample_commands(idx, STREAM_HOVER)
                self.hover_counter[idx] = 0
                self.hover_resampled[idx] = True

        # termination (hover_env.py:481-532)
        cfg = self.env_cfg
        e, rp = self.base_euler, self.rel_pos
        self.crash_condition = (
            (np.abs(e[:, 1]) > F32(cfg["termination_if_pitch_greater_than"]))
            | (np.abs(e[:, 0]) > F32(cfg["termination_if_roll_greater_than"]))
            | (np.abs(rp[:, 0]) > F32(cfg["termination_if_x_greater_than"]))
            | (np.abs(rp[:, 1]) > F32(cfg["termination_if_y_greater_than"]))
            | (np.abs(rp[:, 2]) > F32(cfg["termination_if_z_greater_than"]))
            | (self.base_pos[:, 2] < F32(cfg["termination_if_close_to_ground"]))
            | np.any(np.abs(self.base_ang_vel) > F32(cfg["termination_if_ang_vel_greater_than"]), axis=1)
            | np.any(np.abs(self.base_lin_vel) > F32(cfg["termination_if_lin_vel_greater_than"]), axis=1)
        )
        if cfg["termination_if_close_to_ground"] <= 0.0:
            self.below_ground_env_steps += int((self.base_pos[:, 2] < F32(0.0)).sum())
        timeout = self.episode_length_buf > self.max_episode_length
        self.reset_buf = timeout | self.crash_condition
        self.time_outs = timeout.astype(F32)
        self._reset_idx(np.nonzero(self.reset_buf)[0])

        # rewards (hover_env.py:535-539, 816-860), dict order = summation order
        self.rew_buf = np.zeros(N, F32)
        for name, scale in self.reward_scales.items():
            rew = getattr(self, "_reward_" + name)() * F32(scale)
            self.rew_buf = self.rew_buf + rew
            self.episode_sums[name] = self.episode_sums[name] + rew

        self.obs_buf = self.compute_observations()
        self.last_actions[:] = self.actions
        self.rng.advance()
        return self.obs_buf, self.rew_buf, self.reset_buf, self.time_outs

    # ------------------------------------------------------------------ rewards
    @staticmethod
    def _sumsq(a):
        acc = a[:, 0] * a[:, 0]
        for j in range(1, a.shape[1]):
            acc = acc + a[:, j] * a[:, j]
        return acc

    def _reward_target(self):
        return self._sumsq(self.last_rel_pos) - self._sumsq(self.rel_pos)

    def _reward_closeness(self):
        t2 = F32(self.at_target_threshold_square)
        r = (t2 - self._sumsq(self.rel_pos)) * (F32(1.0) / t2)
        return np.where(r < F32(0.0), F32(0.0), r)

    def _reward_hover_time(self):
        return self.hover_counter.astype(F32)

    def _reward_smooth(self):
        return self._sumsq(self.actions - self.last_actions)

    def _reward_yaw(self):
        yaw = self.base_euler[:, 2]
        yaw = np.where(yaw > F32(180.0), yaw - F32(360.0), yaw) * (F32(1.0) / F32(180.0)) * F32(3.14159)
        return np.exp(F32(self.yaw_lambda) * np.abs(yaw))

    def _reward_angular(self):
        return np.sqrt(self._sumsq(self.base_ang_vel * (F32(1.0) / F32(3.14159))))

    def _reward_crash(self):
        return self.crash_condition.astype(F32)

    # ------------------------------------------------------------------ obs
    def compute_observations(self):
        clip = lambda a: np.where(a < F32(-1.0), F32(-1.0), np.where(a > F32(1.0), F32(1.0), a))  # noqa: E731
        pos = clip(self.rel_pos * F32(self.obs_scales["rel_pos"]))
        quat = self.base_quat.copy()
        lin = clip(self.base_lin_vel * F32(self.obs_scales["lin_vel"]))
        ang = clip(self.base_ang_vel * F32(self.obs_scales["ang_vel"]))
        if self.obs_noise_std > 0.0:
            z = self.rng.normal(np.arange(self.num_envs), STREAM_OBS_NOISE, 13)
            s = F32(self.obs_noise_std)
            pos = pos + np.stack(z[0:3], 1) * s
            lin = lin + np.stack(z[3:6], 1) * s
            ang = ang + np.stack(z[6:9], 1) * s
            quat = quat + np.stack(z[9:13], 1) * s
            n = np.sqrt(((quat[:, 0] ** 2 + quat[:, 1] ** 2) + quat[:, 2] ** 2) + quat[:, 3] ** 2)
            quat = quat / n[:, None]
        return np.concatenate([pos, quat, lin, ang, self.last_actions], axis=1).astype(F32)

    # ------------------------------------------------------------------ target / state
    def update_target_pos(self, idx, target_pos):
        self.commands[np.asarray(idx, dtype=np.int64), :] = np.asarray(target_pos, dtype=F32)

    def get_current_state(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        return OracleEnvState(
            self.base_pos[idx].copy(),
            self.base_quat[idx].copy(),
            self.base_lin_vel[idx].copy(),
            self.base_ang_vel[idx].copy(),
            self.commands[idx].copy(),
            self.episode_length_buf[idx].copy(),
            self.last_actions[idx].copy(),
            self.pid_integral.copy() if self.uses_ctbr else None,
            self.pid_prev_error.copy() if self.uses_ctbr else None,
        )

    def restore_from_state(self, idx, s: OracleEnvState):
        idx = np.asarray(idx, dtype=np.int64)
        self.base_pos[idx] = s.base_pos
        self.last_base_pos[idx] = s.base_pos
        self.rel_pos[idx] = s.commands - s.base_pos
        self.last_rel_pos[idx] = s.commands - s.base_pos
        self.base_quat[idx] = s.base_quat.reshape(1, -1)
        self.sim_pos[idx] = self.base_pos[idx]
        self.sim_quat[idx] = self.base_quat[idx]
        self.base_lin_vel[idx] = s.base_lin_vel
        self.base_ang_vel[idx] = s.base_ang_vel
        # body-frame buffers written into the simulator dofs (hover_env.py:729-735)
        self.sim_vel[idx] = self.base_lin_vel[idx]
        self.sim_ang[idx] = self.base_ang_vel[idx]
        self.last_actions[idx] = s.last_actions
        self.episode_length_buf[idx] = s.episode_length
        self.reset_buf[idx] = True
        self.extras_episode = {}
        for k in self.episode_sums:
            self.extras_episode["rew_" + k] = float(np.mean(self.episode_sums[k][idx])) / self.env_cfg["episode_length_s"]
            self.episode_sums[k][idx] = 0
        if self.uses_ctbr:
            self.pid_integral = s.pid_integral.copy()
            self.pid_prev_error = s.pid_prev_error.copy()
