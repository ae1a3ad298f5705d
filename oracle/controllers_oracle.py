"""TEST INFRASTRUCTURE ONLY (oracle) -- numpy fp32 restatements of the stand-alone controllers.

Follows (reference `src/data_driven_quad_control/`):
  utilities/vectorized_pid_controller.py:97-126     VectorizedPIDController.compute
  controllers/ctbr/ctbr_controller.py:221-264       DroneCTBRController.compute
  controllers/tracking/tracking_controller.py:90-164 DroneTrackingController.compute
  utilities/math_utils.py:23-83                      quaternion_to_matrix, yaw_from_quaternion

Pinned against the reference classes themselves (imported from /root/reference/src in the build
container) by `oracle/make_golden.py` -> `tests/golden/controllers.npz`, checked in
`tests/test_oracle_golden.py`.  Tensor / python-scalar divisions follow torch's CUDA semantics
(x * (1/s)), see oracle/env_oracle.py.
"""

from __future__ import annotations

import numpy as np

F32 = np.float32


class PIDOracle:
    def __init__(self, gains, num_envs, dt):
        g = np.asarray(gains, dtype=F32)
        self.kp, self.ki, self.kd = g[:, 0].copy(), g[:, 1].copy(), g[:, 2].copy()
        self.dt = F32(dt)
        self.integral = np.zeros((num_envs, g.shape[0]), F32)
        self.prev_error = np.zeros((num_envs, g.shape[0]), F32)

    def compute(self, setpoints, measurements):
        err = np.asarray(setpoints, F32) - np.asarray(measurements, F32)
        self.integral = self.integral + err * self.dt
        deriv = (err - self.prev_error) * (F32(1.0) / self.dt)
        self.prev_error = err.copy()
        return (self.kp[None, :] * err + self.ki[None, :] * self.integral) + self.kd[None, :] * deriv

    def reset(self):
        self.integral[:] = 0
        self.prev_error[:] = 0


class CTBROracle:
    def __init__(self, mixer, inv_sqrt_kf, gains, num_envs, dt):
        self.mixer = np.asarray(mixer, F32).reshape(4, 4)
        self.inv_sqrt_kf = F32(inv_sqrt_kf)
        self.pid = PIDOracle(gains, num_envs, dt)

    def compute(self, rate_measurements, rate_setpoints, thrust_setpoints):
        acc = self.pid.compute(rate_setpoints, rate_measurements)
        c = (acc[:, 0], acc[:, 1], acc[:, 2], np.asarray(thrust_setpoints, F32))
        M = self.mixer
        rpm = np.zeros((acc.shape[0], 4), F32)
        for i in range(4):
            f = ((M[i, 0] * c[0] + M[i, 1] * c[1]) + M[i, 2] * c[2]) + M[i, 3] * c[3]
            f = np.where(f < F32(0.0), F32(0.0), f)
            rpm[:, i] = np.sqrt(f) * self.inv_sqrt_kf
        return rpm


def quat_to_matrix(q):
    w, x, y, z = (q[:, j] for j in range(4))
    xx, yy, zz, ww = x * x, y * y, z * z, w * w
    xy, xz, yz, wx, wy, wz = x * y, x * z, y * z, w * x, w * y, w * z
    two = F32(2.0)
    R = np.empty((q.shape[0], 3, 3), F32)
    R[:, 0, 0] = ((ww + xx) - yy) - zz
    R[:, 0, 1] = two * (xy - wz)
    R[:, 0, 2] = two * (xz + wy)
    R[:, 1, 0] = two * (xy + wz)
    R[:, 1, 1] = ((ww - xx) + yy) - zz
    R[:, 1, 2] = two * (yz - wx)
    R[:, 2, 0] = two * (xz - wy)
    R[:, 2, 1] = two * (yz + wx)
    R[:, 2, 2] = ((ww - xx) - yy) + zz
    return R


def _normalize(v):
    n = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])
    n = np.where(n < F32(1e-12), F32(1e-12), n)  # torch.nn.functional.normalize eps
    return v / n[:, None]


def _cross(a, b):
    return np.stack(
        (
            a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
            a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
            a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0],
        ),
        1,
    )


class TrackingOracle:
    def __init__(self, mass, gains, kR, num_envs, dt):
        self.pid = PIDOracle(gains, num_envs, dt)
        self.mass, self.kR = F32(mass), F32(kR)

    def compute(self, X, Q, X_sp, Q_sp):
        X, Q, X_sp, Q_sp = (np.asarray(a, F32) for a in (X, Q, X_sp, Q_sp))
        acc = self.pid.compute(X_sp, X)
        acc[:, 2] = acc[:, 2] + F32(9.81)
        f = acc * self.mass
        R = quat_to_matrix(Q)
        thrust = (R[:, 2, 0] * f[:, 0] + R[:, 2, 1] * f[:, 1]) + R[:, 2, 2] * f[:, 2]
        z_des = _normalize(f)
        w, x, y, z = (Q_sp[:, j] for j in range(4))
        yaw = np.arctan2(F32(2.0) * (w * z + x * y), F32(1.0) - F32(2.0) * (y * y + z * z))
        x_yaw = np.stack((np.cos(yaw), np.sin(yaw), np.zeros_like(yaw)), 1).astype(F32)
        y_des = _normalize(_cross(z_des, x_yaw))
        x_des = _normalize(_cross(y_des, z_des))
        Rd = np.stack((x_des, y_des, z_des), 2)  # columns
        M = np.empty_like(R)
        for a in range(3):
            for b in range(3):
                M[:, a, b] = (Rd[:, 0, a] * R[:, 0, b] + Rd[:, 1, a] * R[:, 1, b]) + Rd[:, 2, a] * R[:, 2, b]
        e = np.stack(
            (F32(0.5) * (M[:, 2, 1] - M[:, 1, 2]), F32(0.5) * (M[:, 0, 2] - M[:, 2, 0]), F32(0.5) * (M[:, 1, 0] - M[:, 0, 1])),
            1,
        )
        return np.concatenate((thrust[:, None], -self.kR * e), 1).astype(F32)
