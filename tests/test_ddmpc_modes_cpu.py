"""CPU: the dual-recovery algebra behind alpha_reg_type PREVIOUS and update_cost_threshold (numpy mirror of
csrc/ddqc_ddmpc_recover.cu, oracle/ddmpc_condensed.py::recover_dual) against the literal-QP oracle."""

import numpy as np
import pytest

import ddmpc_condensed as dc
import ddmpc_oracle as mo


def _problem(seed):
    prm = mo.default_params()
    prm.n, prm.L, prm.N = 2, 5, 80
    rng = np.random.default_rng(seed)
    A = np.diag(rng.uniform(0.6, 0.95, 6)) + 0.02 * rng.standard_normal((6, 6))
    Bm, Cm = rng.standard_normal((6, 3)) * 0.3, rng.standard_normal((3, 6)) * 0.5
    x = np.zeros(6)

    def step(u):
        nonlocal x
        x = A @ x + Bm @ u
        return Cm @ x + 1e-4 * rng.standard_normal(3)

    u = np.stack([rng.uniform(0.15, 0.4, prm.N), rng.uniform(-0.5, 0.5, prm.N), rng.uniform(-0.5, 0.5, prm.N)], 1)
    y = np.stack([step(u[k]) for k in range(prm.N)])
    return prm, u, y, step


@pytest.mark.parametrize("alpha_reg_type,thr", [(1, None), (0, 9.0), (2, 57.0)])
def test_recovered_alpha_and_tracking_cost_match_the_literal_qp(alpha_reg_type, thr):
    prm, u, y, step = _problem(5)
    prm.alpha_reg_type = alpha_reg_type
    y_r = np.array([0.3, -0.2, 0.5])
    orc = mo.NonlinearDDMPCOracle(**{**mo.ctor_kwargs(prm, u, y, y_r), "update_cost_threshold": thr}, solve_on_init=False)
    ur, yr = u.copy(), y.copy()
    ud, yd = ur.copy(), yr.copy()
    alpha_prev, cost, froze = np.zeros(prm.C), None, 0
    for t in range(5):
        orc.update_and_solve_data_driven_mpc()
        if thr is None or cost is None or cost > thr:
            ud, yd = ur.copy(), yr.copy()
        else:
            froze += 1
        _, H = dc.hankel_gram(ud, yd, prm.ell)
        v_sr = H @ alpha_prev if alpha_reg_type == 1 else None
        s = dc.solve_condensed(prm, ud, yd, y_r, v_sr_ext=v_sr, u_ic=ur, y_ic=yr)
        _, Htz, cost = dc.recover_dual(prm, ud, yd, y_r, s["u_opt"], s["us"], s["ys"], v_sr=s["v_sr"], u_ic=ur, y_ic=yr)
        np.testing.assert_allclose(s["u_opt"], orc.optimal_u, rtol=0, atol=1e-6)
        assert abs(cost - orc.tracking_cost) <= 1e-7 * max(1.0, cost)
        if alpha_reg_type == 1:
            alpha_prev = alpha_prev + Htz
            np.testing.assert_allclose(alpha_prev, orc.alpha_prev, rtol=0, atol=1e-6)
        u0 = orc.optimal_u[0]
        yn = step(u0)
        orc.store_input_output_measurement(u0, yn)
        ur, yr = np.vstack([ur[1:], u0]), np.vstack([yr[1:], yn])
    if thr is not None:
        assert froze > 0
