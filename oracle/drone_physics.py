"""TEST INFRASTRUCTURE ONLY (oracle) -- restated Genesis drone rigid-body substep.

PARITY UNPINNED for this file: the arithmetic of the physics substep lives in
the un-vendored third-party package `genesis_world @ git+.../Genesis@1249131`
(reference `pyproject.toml:15`), which is neither under /root/reference nor
installed here.  The reference only pins it *behaviourally*
(`tests/controller_tests/test_drone_ctbr_controller.py:45-104`: hover |w|<0.1
after 10 steps, rate error <= 1e-3 after 20 steps), both of which this
restatement satisfies (see tests/test_oracle_reference_parity.py).

Restated model (call sites: reference `hover_env.py:437-439` set_propellels_rpm +
scene.step(); `hover_env.py:452-470` get_pos/get_quat/get_vel/get_ang):

  per rotor i:   T_i = (rpm_i*rpm_i)*KF, body force [0,0,T_i] at r_i=(x_i,y_i,0)
                 yaw torque  sum_i ((rpm_i*rpm_i)*KM)*spin_i  (body z)
  body rates:    J wdot = tau - w x (J w)            (free-joint angular dofs in the body frame)
  world linear:  m vdot = R(q) [0,0,sum T] + m g     (g = -9.81 z)
  integrator:    semi-implicit Euler, dt = sim dt:  w,v first (forces at the old pose),
                 then p += v*dt and q <- normalize(q (x) dq), dq = [cos|h|, sinc|h| h], h = w*dt/2,
                 with cos / sinc evaluated by their Taylor series in |h|^2 (exact to fp32
                 rounding for |h| <= 1 rad, i.e. |w| <= 200 rad/s at dt = 0.01).
  no ground contact (default config terminates at z < 0.1 m).

Every function here uses ONLY elementwise +,-,*,/ and sqrt in a fixed order, so
it gives bit-identical fp32 results on torch-CPU tensors (the genesis shim) and
on numpy float32 arrays (the env oracle), and can be mirrored bit-exactly by a
CUDA kernel compiled without FMA contraction.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# cf2x.urdf values (Genesis asset `urdf/drones/cf2x.urdf`, mirrored by reference
# `configs/drone/cf2x_drone_params.yaml:10-30`)
CF2X_MASS = 0.027
CF2X_INERTIA = (1.4e-5, 1.4e-5, 2.17e-5)
CF2X_KF = 3.16e-10
CF2X_KM = 7.94e-12
# prop{0..3}_link origins in the URDF (x, y); z = 0
CF2X_ROTOR_XY = ((0.028, -0.028), (-0.028, -0.028), (-0.028, 0.028), (0.028, 0.028))
CF2X_SPIN = (-1.0, 1.0, -1.0, 1.0)
GRAVITY = 9.81

# Taylor coefficients of cos(a) and sin(a)/a as polynomials in s = a*a
COS_COEF = tuple((-1.0) ** k / math.factorial(2 * k) for k in range(7))
SINC_COEF = tuple((-1.0) ** k / math.factorial(2 * k + 1) for k in range(7))


@dataclass(frozen=True)
class PhysParams:
    """fp32-rounded constants shared by oracle, shim and (via the C struct) the kernel."""

    dt: np.float32
    half_dt: np.float32
    kf: np.float32
    km: np.float32
    inv_mass: np.float32
    gravity: np.float32
    jxx: np.float32
    jyy: np.float32
    jzz: np.float32
    inv_jxx: np.float32
    inv_jyy: np.float32
    inv_jzz: np.float32
    rotor_x: tuple
    rotor_y: tuple
    spin: tuple
    cos_coef: tuple
    sinc_coef: tuple

    @staticmethod
    def cf2x(dt: float) -> "PhysParams":
        f = np.float32
        return PhysParams(
            dt=f(dt),
            half_dt=f(dt / 2.0),
            kf=f(CF2X_KF),
            km=f(CF2X_KM),
            inv_mass=f(1.0 / CF2X_MASS),
            gravity=f(GRAVITY),
            jxx=f(CF2X_INERTIA[0]),
            jyy=f(CF2X_INERTIA[1]),
            jzz=f(CF2X_INERTIA[2]),
            inv_jxx=f(1.0 / CF2X_INERTIA[0]),
            inv_jyy=f(1.0 / CF2X_INERTIA[1]),
            inv_jzz=f(1.0 / CF2X_INERTIA[2]),
            rotor_x=tuple(f(x) for x, _ in CF2X_ROTOR_XY),
            rotor_y=tuple(f(y) for _, y in CF2X_ROTOR_XY),
            spin=tuple(f(s) for s in CF2X_SPIN),
            cos_coef=tuple(f(c) for c in COS_COEF),
            sinc_coef=tuple(f(c) for c in SINC_COEF),
        )


def _c(x):
    """fp32 scalar -> python float (exactly representable), so that the same
    expression works on torch tensors and numpy arrays without dtype promotion."""
    return float(x)


def _horner(coef, s):
    acc = s * _c(coef[6]) + _c(coef[5])
    for k in (4, 3, 2, 1, 0):
        acc = acc * s + _c(coef[k])
    return acc


def quat_mul(aw, ax, ay, az, bw, bx, by, bz):
    """Hamilton product a (x) b, (w, x, y, z)."""
    w = ((aw * bw - ax * bx) - ay * by) - az * bz
    x = ((aw * bx + ax * bw) + ay * bz) - az * by
    y = ((aw * by - ax * bz) + ay * bw) + az * bx
    z = ((aw * bz + ax * by) - ay * bx) + az * bw
    return w, x, y, z


def rotate_by_quat(vx, vy, vz, qw, qx, qy, qz):
    """Genesis `genesis.utils.geom.transform_by_quat(v, q)`: v + w t + qv x t, t = 2 qv x v."""
    tx = (qy * vz - qz * vy) * 2.0
    ty = (qz * vx - qx * vz) * 2.0
    tz = (qx * vy - qy * vx) * 2.0
    rx = (vx + qw * tx) + (qy * tz - qz * ty)
    ry = (vy + qw * ty) + (qz * tx - qx * tz)
    rz = (vz + qw * tz) + (qx * ty - qy * tx)
    return rx, ry, rz


def substep(P: PhysParams, sqrt, p, q, v, w, rpm):
    """One rigid-body substep.  p,v,w: 3-tuples; q: 4-tuple (w,x,y,z); rpm: 4-tuple.
    `sqrt` is the elementwise square root of the array library in use."""
    px, py, pz = p
    qw, qx, qy, qz = q
    vx, vy, vz = v
    wx, wy, wz = w

    # rotor thrusts and body torques
    t = [(r * r) * _c(P.kf) for r in rpm]
    fz = ((t[0] + t[1]) + t[2]) + t[3]
    tau_x = ((t[0] * _c(P.rotor_y[0]) + t[1] * _c(P.rotor_y[1])) + t[2] * _c(P.rotor_y[2])) + t[3] * _c(P.rotor_y[3])
    tau_y = ((t[0] * _c(-P.rotor_x[0]) + t[1] * _c(-P.rotor_x[1])) + t[2] * _c(-P.rotor_x[2])) + t[3] * _c(-P.rotor_x[3])
    m = [((r * r) * _c(P.km)) * _c(s) for r, s in zip(rpm, P.spin)]
    tau_z = ((m[0] + m[1]) + m[2]) + m[3]

    # body-frame Euler equations
    jwx = wx * _c(P.jxx)
    jwy = wy * _c(P.jyy)
    jwz = wz * _c(P.jzz)
    gx = wy * jwz - wz * jwy
    gy = wz * jwx - wx * jwz
    gz = wx * jwy - wy * jwx
    wx = wx + ((tau_x - gx) * _c(P.inv_jxx)) * _c(P.dt)
    wy = wy + ((tau_y - gy) * _c(P.inv_jyy)) * _c(P.dt)
    wz = wz + ((tau_z - gz) * _c(P.inv_jzz)) * _c(P.dt)

    # world-frame linear acceleration: third column of R(q) times thrust / m, plus gravity
    r13 = (qx * qz + qw * qy) * 2.0
    r23 = (qy * qz - qw * qx) * 2.0
    r33 = ((qw * qw - qx * qx) - qy * qy) + qz * qz
    acc = fz * _c(P.inv_mass)
    vx = vx + (r13 * acc) * _c(P.dt)
    vy = vy + (r23 * acc) * _c(P.dt)
    vz = vz + (r33 * acc - _c(P.gravity)) * _c(P.dt)

    # positions with the NEW velocities (semi-implicit Euler)
    px = px + vx * _c(P.dt)
    py = py + vy * _c(P.dt)
    pz = pz + vz * _c(P.dt)

    hx = wx * _c(P.half_dt)
    hy = wy * _c(P.half_dt)
    hz = wz * _c(P.half_dt)
    s = (hx * hx + hy * hy) + hz * hz
    c = _horner(P.cos_coef, s)
    k = _horner(P.sinc_coef, s)
    nw, nx, ny, nz = quat_mul(qw, qx, qy, qz, c, k * hx, k * hy, k * hz)
    inv_n = 1.0 / sqrt(((nw * nw + nx * nx) + ny * ny) + nz * nz)
    q = (nw * inv_n, nx * inv_n, ny * inv_n, nz * inv_n)

    return (px, py, pz), q, (vx, vy, vz), (wx, wy, wz)
