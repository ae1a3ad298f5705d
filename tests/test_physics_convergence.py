"""What CAN be pinned about the restated rigid-body substep without Genesis (oracle/drone_physics.py, PARITY UNPINNED):
it is a consistent, first-order convergent discretisation of the Newton-Euler equations of the cf2x airframe.

Genesis integrates the same continuous model (free joint, semi-implicit Euler, external rotor forces); two consistent
first-order schemes agree up to O(dt) per unit time, which is the size of disagreement the unpinned conventions
(gyroscopic term placement, quaternion update form) can cause - the test measures that size.  Reference solution: the
continuous dynamics in float64 integrated by scipy's DOP853 at rtol 1e-12."""

import numpy as np
import pytest
from scipy.integrate import solve_ivp

import drone_physics as dp

F32 = np.float32


def _quat_rot(q, v):
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])  # fmt: skip
    return R @ v


def _rhs(rpm):
    J = np.array(dp.CF2X_INERTIA)
    T = dp.CF2X_KF * rpm**2
    tau = np.zeros(3)
    for (x, y), t in zip(dp.CF2X_ROTOR_XY, T):
        tau += np.cross([x, y, 0.0], [0.0, 0.0, t])
    tau[2] += float(np.sum(dp.CF2X_KM * np.array(dp.CF2X_SPIN) * rpm**2))
    thrust = np.array([0.0, 0.0, T.sum()])

    def f(t, s):
        p, q, v, w = s[0:3], s[3:7], s[7:10], s[10:13]
        qn = q / np.linalg.norm(q)
        dv = _quat_rot(qn, thrust) / dp.CF2X_MASS + np.array([0.0, 0.0, -dp.GRAVITY])
        dw = (tau - np.cross(w, J * w)) / J
        qw, qx, qy, qz = qn
        dq = 0.5 * np.array([-qx * w[0] - qy * w[1] - qz * w[2], qw * w[0] + qy * w[2] - qz * w[1],
                             qw * w[1] - qx * w[2] + qz * w[0], qw * w[2] + qx * w[1] - qy * w[0]])  # fmt: skip
        return np.concatenate([v, dq, dv, dw])

    return f


def _oracle_run(dt, steps, rpm, s0):
    P = dp.PhysParams.cf2x(dt)
    wrench = dp.rotor_wrench(P, [F32(r) for r in rpm])
    p = tuple(F32(x) for x in s0[0:3])
    q = tuple(F32(x) for x in s0[3:7])
    v = tuple(F32(x) for x in s0[7:10])
    w = tuple(F32(x) for x in s0[10:13])
    for _ in range(steps):
        p, q, v, w = dp.substep(P, wrench, p, q, v, w)
    return np.array(p, np.float64), np.array(q, np.float64), np.array(v, np.float64), np.array(w, np.float64)


S0 = np.array([0.0, 0.0, 1.0, 0.9900, 0.0800, -0.0600, 0.0950, 0.3, -0.2, 0.1, 0.8, -0.5, 0.4])
S0[3:7] /= np.linalg.norm(S0[3:7])
RPM = np.array([14000.0, 14650.0, 15100.0, 14300.0])


def test_substep_converges_at_first_order_to_the_newton_euler_solution():
    T = 0.4  # ten env steps of 4 substeps at the reference's dt = 0.01
    sol = solve_ivp(_rhs(RPM), (0.0, T), S0, method="DOP853", rtol=1e-12, atol=1e-14)
    ref = sol.y[:, -1]
    ref_q = ref[3:7] / np.linalg.norm(ref[3:7])
    errs = []
    for dt in (0.01, 0.005, 0.0025):
        p, q, v, w = _oracle_run(dt, int(round(T / dt)), RPM, S0)
        if np.dot(q, ref_q) < 0:
            q = -q
        errs.append((np.abs(p - ref[0:3]).max(), np.abs(q - ref_q).max(), np.abs(v - ref[7:10]).max(), np.abs(w - ref[10:13]).max()))
    errs = np.array(errs)
    # first order: halving dt halves the error of every state component (fp32 round-off is far below these errors)
    for k in range(4):
        assert 1.6 < errs[0, k] / errs[1, k] < 2.4 and 1.6 < errs[1, k] / errs[2, k] < 2.4, errs[:, k]
    # size of the discretisation error at the reference's dt over 0.4 s of an aggressive manoeuvre (rates growing past
    # 10 rad/s): measured 8.8 mm in position, 2.7e-2 in the quaternion, 5.4e-3 m/s, 8.0e-3 rad/s - this O(dt) term is also
    # the most that two consistent semi-implicit schemes (this one and Genesis') can differ by over such a manoeuvre
    assert errs[0, 0] < 2e-2 and errs[0, 1] < 5e-2, errs[0]


def test_hover_is_an_equilibrium_and_free_fall_is_exact():
    """Zero rates, level attitude, rotor speed sqrt(m g / (4 KF)): the drone stays put up to fp32 rounding of the torque
    sums (4 s: 0.5 mm of drift, 2e-5 rad/s; the reference's own pin is |w| < 0.1 after 10 env steps).  Zero thrust:
    v_z = -g t exactly, z = z0 - g dt^2 k (k - 1) / 2 in the pose-first order (the pose advances with the velocity of the
    previous substep; k (k + 1) / 2 in the velocity-first order used until round 2)."""
    hover_rpm = np.sqrt(dp.CF2X_MASS * dp.GRAVITY / (4 * dp.CF2X_KF))
    s0 = np.array([0.0, 0.0, 1.0, 1.0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    p, q, v, w = _oracle_run(0.01, 400, np.full(4, hover_rpm), s0)
    assert np.abs(p - [0, 0, 1]).max() < 2e-3 and np.abs(v).max() < 2e-3 and np.abs(w).max() < 1e-4 and abs(q[0] - 1) < 1e-7
    k = 50
    p, q, v, w = _oracle_run(0.01, k, np.zeros(4), s0)
    assert abs(v[2] + dp.GRAVITY * 0.01 * k) < 1e-5 and abs(p[2] - (1.0 - dp.GRAVITY * 1e-4 * k * (k - 1) / 2)) < 1e-5


def test_quaternion_stays_normalised_and_yaw_equivariance():
    """|q| = 1 to fp32 rounding over 1000 substeps of a tumbling body; rotating the initial state about the vertical axis
    rotates the trajectory (the model has no preferred heading)."""
    p, q, v, w = _oracle_run(0.01, 1000, RPM, S0)
    assert abs(np.linalg.norm(q) - 1.0) < 5e-7
    psi = 0.7
    rz = np.array([np.cos(psi / 2), 0.0, 0.0, np.sin(psi / 2)])

    def qmul(a, b):
        return np.array(dp.quat_mul(*[F32(x) for x in a], *[F32(x) for x in b]), np.float64)

    Rz = np.array([[np.cos(psi), -np.sin(psi), 0], [np.sin(psi), np.cos(psi), 0], [0, 0, 1]])
    s1 = S0.copy()
    s1[0:3], s1[7:10] = Rz @ S0[0:3], Rz @ S0[7:10]
    s1[3:7] = qmul(rz, S0[3:7])  # world-frame yaw applied on the left; body rates unchanged
    pa, qa, va, wa = _oracle_run(0.01, 40, RPM, S0)
    pb, qb, vb, wb = _oracle_run(0.01, 40, RPM, s1)
    np.testing.assert_allclose(pb, Rz @ pa, atol=2e-6)
    np.testing.assert_allclose(vb, Rz @ va, atol=2e-5)
    np.testing.assert_allclose(wb, wa, atol=2e-5)
    qe = qmul(rz, qa)
    assert min(np.abs(qb - qe).max(), np.abs(qb + qe).max()) < 2e-6
