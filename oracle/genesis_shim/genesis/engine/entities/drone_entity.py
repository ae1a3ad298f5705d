"""TEST INFRASTRUCTURE ONLY -- restated Genesis `DroneEntity` for a single free body.

State per env: pos (world), quat (w,x,y,z), linear velocity (world), angular
velocity (free-joint angular dofs, body frame).  `scene.step()` advances it with
`oracle/drone_physics.substep`; external forces are cleared each step, so
`set_propellels_rpm` must be called before every step (as the reference does,
`hover_env.py:437-439`).
"""

import numpy as np
import torch

import drone_physics as _phys

from .rigid_entity import RigidEntity


class DroneEntity(RigidEntity):
    def _build(self, n_envs, dt):
        self.n_envs = n_envs
        self.P = _phys.PhysParams.cf2x(dt)
        z = lambda k: torch.zeros((n_envs, k), dtype=torch.float32)  # noqa: E731
        self.pos, self.quat, self.vel, self.ang = z(3), z(4), z(3), z(3)
        self.quat[:, 0] = 1.0
        self.rpm = z(4)

    # --- actuation / stepping -------------------------------------------------
    def set_propellels_rpm(self, rpm):
        rpm = torch.as_tensor(rpm, dtype=torch.float32)
        if (rpm < 0).any():
            raise ValueError("rpm must be non-negative")
        self.rpm = rpm.reshape(self.n_envs, 4).clone()

    @staticmethod
    def _np_cols(t):
        a = t.numpy()
        return tuple(a[:, j].copy() for j in range(a.shape[1]))

    def _step(self):
        wrench = _phys.rotor_wrench(self.P, self._np_cols(self.rpm))
        p, q, v, w = _phys.substep(
            self.P, wrench, self._np_cols(self.pos), self._np_cols(self.quat), self._np_cols(self.vel),
            self._np_cols(self.ang),
        )
        stack = lambda cols: torch.from_numpy(np.stack(cols, 1).astype(np.float32))  # noqa: E731
        self.pos, self.quat, self.vel, self.ang = stack(p), stack(q), stack(v), stack(w)
        self.rpm = torch.zeros_like(self.rpm)  # external forces are cleared every step

    # --- getters (world frame) --------------------------------------------------
    def get_pos(self):
        return self.pos.clone()

    def get_quat(self):
        return self.quat.clone()

    def get_vel(self):
        return self.vel.clone()

    def get_ang(self):
        # cd_ang of the base link = R(q) w_body
        cols = _phys.rotate_by_quat(*self._np_cols(self.ang), *self._np_cols(self.quat))
        return torch.from_numpy(np.stack(cols, 1).astype(np.float32))

    # --- setters ----------------------------------------------------------------
    @staticmethod
    def _idx(envs_idx):
        if isinstance(envs_idx, (list, tuple)) and len(envs_idx) == 1 and torch.is_tensor(envs_idx[0]):
            envs_idx = envs_idx[0]  # reference passes `envs_idx=[tensor]` at hover_env.py:735
        return torch.as_tensor(envs_idx, dtype=torch.long).reshape(-1)

    def set_pos(self, pos, zero_velocity=True, envs_idx=None):
        idx = self._idx(envs_idx)
        self.pos[idx] = torch.as_tensor(pos, dtype=torch.float32).reshape(-1, 3)
        if zero_velocity:
            self.zero_all_dofs_velocity(idx)

    def set_quat(self, quat, zero_velocity=True, envs_idx=None):
        idx = self._idx(envs_idx)
        self.quat[idx] = torch.as_tensor(quat, dtype=torch.float32).reshape(-1, 4)
        if zero_velocity:
            self.zero_all_dofs_velocity(idx)

    def zero_all_dofs_velocity(self, envs_idx=None):
        idx = self._idx(envs_idx)
        self.vel[idx] = 0.0
        self.ang[idx] = 0.0

    def set_dofs_velocity(self, velocity, envs_idx=None):
        idx = self._idx(envs_idx)
        velocity = torch.as_tensor(velocity, dtype=torch.float32).reshape(-1, 6)
        self.vel[idx] = velocity[:, :3]
        self.ang[idx] = velocity[:, 3:]
