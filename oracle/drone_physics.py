"""TEST INFRASTRUCTURE ONLY (oracle) -- restated Genesis drone rigid-body substep.

The arithmetic of the physics substep lives in the un-vendored third-party package
`genesis_world @ git+.../Genesis@1249131` (reference `pyproject.toml:15`), which is neither under
/root/reference nor installable here (profiles/r2_reference_stack_probe.log), so this file cannot be
compared with Genesis call by call.  What pins it:

  * the ONE output of the real stack that ships with the reference, `docs/resources/
    controller_comparison_plot.png` (Genesis + the reference's tracking controller / PPO policy), digitised by
    `oracle/digitize_reference_plot.py` into `tests/golden/ref_comparison_plot_curves.npz`:
    `tests/test_reference_plot_pin.py` flies the same scenario on this restatement and requires the
    tracking-controller drone and the PPO drone to stay within a few pixels of the published curves over all
    14 s (1 px = 0.047 s, ~6 mm), and the PPO drone to hover as steadily as the reference's does.  That
    figure selected the ORDER of the update below: with the velocity-first (semi-implicit) order used until
    round 2 the same flights missed the published curves by up to 12.6 px (tracking overshoot 0.10 m instead
    of 0.17 m at the second setpoint) and the PPO policy showed a 20 mm limit cycle in hover that the
    reference's run does not have; pose-first reproduces both (max 2.1 / 4.9 px, 0.2 mm).  Candidate orders
    compared on the figure (sum of the six 90th-percentile pixel distances): velocity-first 17.6, fully
    explicit 10.4, pose-first 9.8, velocity-first with the rotor command delayed one substep 9.0 (needs extra
    per-env state; not adopted).  Pose-first is what a simulator shows when the pose it reports lags its
    generalised coordinates by one substep.
  * behaviourally, the reference's own tests (`tests/controller_tests/test_drone_ctbr_controller.py:45-104`:
    hover |w|<0.1 after 10 steps, rate error <= 1e-3 after 20 steps) pass on it
    (tests/test_oracle_vs_reference.py), and tests/test_physics_convergence.py shows first-order convergence
    to the continuous Newton-Euler solution.

Restated model (call sites: reference `hover_env.py:437-439` set_propellels_rpm +
scene.step(); `hover_env.py:452-470` get_pos/get_quat/get_vel/get_ang):

  per rotor i:   T_i = (rpm_i*rpm_i)*KF, body force [0,0,T_i] at r_i=(x_i,y_i,0)
                 yaw torque  sum_i (rpm_i*rpm_i)*(KM*spin_i)  (body z)
  body rates:    J wdot = tau - w x (J w), J diagonal  (free-joint angular dofs in the body frame)
  world linear:  m vdot = R(q) [0,0,sum T] + m g       (g = -9.81 z)
  integrator:    dt = sim dt, POSE FIRST: p += v*dt and q <- q (x) dq renormalised with the OLD v, w
                 (dq = [cos|h|, sinc|h| h], h = w*dt/2), then w, v with the forces at the NEW pose;
                 cos / sinc by their degree-4 Taylor polynomials in |h|^2 (1 ulp for |w| <= 100 rad/s
                 at dt = 0.01) and the renormalisation by one Newton step q*(1.5 - 0.5|q|^2)
                 (|q|^2 = 1 + O(1e-7) every step, so this equals q/|q| to fp32 rounding).
  no ground contact (default config terminates at z < 0.1 m).

Arithmetic contract.  Every function is a fixed sequence of IEEE fp32 operations: `*`, `+`,
`-` (separately rounded) and `fma(a, b, c)` = round(a*b + c) (single rounding).  The CUDA
kernel (compiled with -fmad=false, explicit __fmaf_rn) executes the same sequence, so it
agrees with this file bit for bit.  `fma` is emulated exactly in float64 (exact product,
TwoSum error term, round-to-odd, then one rounding to float32).
"""

from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np

F32 = np.float32
F64 = np.float64

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

N_TAYLOR = 5
COS_COEF = tuple((-1.0) ** k / math.factorial(2 * k) for k in range(N_TAYLOR))
SINC_COEF = tuple((-1.0) ** k / math.factorial(2 * k + 1) for k in range(N_TAYLOR))


# Timing-only switch (bench.py cpu_baseline): a CPU user would not emulate FMAs in software, so the
# baseline run evaluates fma(a,b,c) as the two-rounding a*b+c.  Never set in the parity tests.
FAST_FMA = os.environ.get("DDQC_ORACLE_FAST_FMA", "0") == "1"


def fma(a, b, c):
    """Correctly rounded float32 fused multiply-add of float32 arrays / scalars."""
    if FAST_FMA:
        return np.asarray(a, dtype=F32) * np.asarray(b, dtype=F32) + np.asarray(c, dtype=F32)
    a64 = np.asarray(a, dtype=F32).astype(F64)
    b64 = np.asarray(b, dtype=F32).astype(F64)
    c64 = np.asarray(c, dtype=F32).astype(F64)
    p = a64 * b64  # exact: 24 + 24 significand bits
    t = p + c64  # round-to-nearest float64
    bb = t - p
    err = (p - (t - bb)) + (c64 - bb)  # exact error of the addition (TwoSum)
    t = np.atleast_1d(t).copy()
    err = np.broadcast_to(err, t.shape)
    ti = t.view(np.int64)
    fix = (err != 0) & np.isfinite(t) & ((ti & 1) == 0)
    up = (err > 0) == (t > 0)  # exact value lies further from zero than t
    ti[fix & up] += 1  # round-to-odd: sticky information survives the final rounding
    ti[fix & ~up] -= 1
    out = t.astype(F32)
    return out if out.shape != (1,) or np.ndim(a) or np.ndim(b) or np.ndim(c) else out[0]


@dataclass(frozen=True)
class PhysParams:
    """fp32 constants, each the rounding of the double-precision expression given here; the
    product's host code (`envs/hover_env.py: make_env_params`) evaluates the same expressions."""

    dt: np.float32
    half_dt: np.float32
    dt_g: np.float32  # dt * 9.81
    kf: np.float32
    dt_over_m: np.float32  # dt / mass
    two_dt_over_m: np.float32  # 2 dt / mass
    k_w: tuple  # dt / J_ii
    b_w: tuple  # dt (J_kk - J_jj) / J_ii  (i,j,k cyclic)
    arm_x: tuple  # roll-torque arm of rotor i:  y_i
    arm_y: tuple  # pitch-torque arm of rotor i: -x_i
    km_spin: tuple  # KM * spin_i
    cos_coef: tuple
    sinc_coef: tuple

    @staticmethod
    def cf2x(dt: float) -> "PhysParams":
        f = np.float32
        jx, jy, jz = CF2X_INERTIA
        return PhysParams(
            dt=f(dt),
            half_dt=f(dt / 2.0),
            dt_g=f(dt * GRAVITY),
            kf=f(CF2X_KF),
            dt_over_m=f(dt / CF2X_MASS),
            two_dt_over_m=f(2.0 * dt / CF2X_MASS),
            k_w=(f(dt / jx), f(dt / jy), f(dt / jz)),
            b_w=(f(dt * (jz - jy) / jx), f(dt * (jx - jz) / jy), f(dt * (jy - jx) / jz)),
            arm_x=tuple(f(y) for _, y in CF2X_ROTOR_XY),
            arm_y=tuple(f(-x) for x, _ in CF2X_ROTOR_XY),
            km_spin=tuple(f(CF2X_KM * s) for s in CF2X_SPIN),
            cos_coef=tuple(f(c) for c in COS_COEF),
            sinc_coef=tuple(f(c) for c in SINC_COEF),
        )


def _horner(coef, s):
    acc = fma(s, coef[4], coef[3])
    acc = fma(acc, s, coef[2])
    acc = fma(acc, s, coef[1])
    return fma(acc, s, coef[0])


def quat_mul(aw, ax, ay, az, bw, bx, by, bz):
    """Hamilton product a (x) b, (w, x, y, z)."""
    w = fma(-az, bz, fma(-ay, by, fma(-ax, bx, aw * bw)))
    x = fma(-az, by, fma(ay, bz, fma(ax, bw, aw * bx)))
    y = fma(az, bx, fma(ay, bw, fma(-ax, bz, aw * by)))
    z = fma(az, bw, fma(-ay, bx, fma(ax, by, aw * bz)))
    return w, x, y, z


def rotate_by_quat(vx, vy, vz, qw, qx, qy, qz):
    """Genesis `genesis.utils.geom.transform_by_quat(v, q)`: v + 2 (w u + qv x u), u = qv x v."""
    ux = fma(qy, vz, -(qz * vy))
    uy = fma(qz, vx, -(qx * vz))
    uz = fma(qx, vy, -(qy * vx))
    sx = fma(qy, uz, fma(-qz, uy, qw * ux))
    sy = fma(qz, ux, fma(-qx, uz, qw * uy))
    sz = fma(qx, uy, fma(-qy, ux, qw * uz))
    two = F32(2.0)
    return fma(two, sx, vx), fma(two, sy, vy), fma(two, sz, vz)


def rotor_wrench(P: PhysParams, rpm):
    """Per-step constants of the zero-order-held rotor command: angular-velocity increments
    from the torques (a_x, a_y, a_z) and the thrust factors of the linear update."""
    sq = [r * r for r in rpm]
    t = [s * P.kf for s in sq]
    fz = ((t[0] + t[1]) + t[2]) + t[3]
    tau_x = fma(t[3], P.arm_x[3], fma(t[2], P.arm_x[2], fma(t[1], P.arm_x[1], t[0] * P.arm_x[0])))
    tau_y = fma(t[3], P.arm_y[3], fma(t[2], P.arm_y[2], fma(t[1], P.arm_y[1], t[0] * P.arm_y[0])))
    tau_z = fma(sq[3], P.km_spin[3], fma(sq[2], P.km_spin[2], fma(sq[1], P.km_spin[1], sq[0] * P.km_spin[0])))
    return (P.k_w[0] * tau_x, P.k_w[1] * tau_y, P.k_w[2] * tau_z, P.two_dt_over_m * fz, P.dt_over_m * fz)


def substep(P: PhysParams, wrench, p, q, v, w):
    """One rigid-body substep.  p,v,w: 3-tuples; q: 4-tuple (w,x,y,z); wrench from rotor_wrench."""
    a_x, a_y, a_z, c_xy, c_z = wrench
    px, py, pz = p
    qw, qx, qy, qz = q
    vx, vy, vz = v
    wx, wy, wz = w

    # pose first, with the OLD velocities (see the module docstring: the order pinned by the reference's published plot)
    px = fma(vx, P.dt, px)
    py = fma(vy, P.dt, py)
    pz = fma(vz, P.dt, pz)

    hx = wx * P.half_dt
    hy = wy * P.half_dt
    hz = wz * P.half_dt
    s = fma(hz, hz, fma(hy, hy, hx * hx))
    c = _horner(P.cos_coef, s)
    k = _horner(P.sinc_coef, s)
    nw, nx, ny, nz = quat_mul(qw, qx, qy, qz, c, k * hx, k * hy, k * hz)
    n2 = fma(nz, nz, fma(ny, ny, fma(nx, nx, nw * nw)))
    inv_n = fma(F32(-0.5), n2, F32(1.5))
    qw, qx, qy, qz = nw * inv_n, nx * inv_n, ny * inv_n, nz * inv_n
    q = (qw, qx, qy, qz)

    # body-frame Euler equations (gyroscopic term with the OLD rates)
    nwx = fma(-P.b_w[0], wy * wz, wx + a_x)
    nwy = fma(-P.b_w[1], wz * wx, wy + a_y)
    nwz = fma(-P.b_w[2], wx * wy, wz + a_z)
    wx, wy, wz = nwx, nwy, nwz

    # world-frame linear velocity: third column of R(q) at the NEW pose times thrust/m, plus gravity
    vx = fma(c_xy, fma(qx, qz, qw * qy), vx)
    vy = fma(c_xy, fma(qy, qz, -(qw * qx)), vy)
    r33 = fma(F32(-2.0), fma(qx, qx, qy * qy), F32(1.0))
    vz = fma(c_z, r33, vz - P.dt_g)

    return (px, py, pz), q, (vx, vy, vz), (wx, wy, wz)
