"""The OSQP-style ADMM port (oracle/osqp_port.py: the CPU baseline of the reference's cvxpy + OSQP class and the yardstick
for "how far is a 1e-5 first-order solution from the exact optimum"): correctness on QPs with known solutions, and its
distance to the high-accuracy literal-QP oracle on the DD-MPC problems."""

import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_equality_constrained_qp_matches_kkt_solution():
    import osqp_port as op

    rng = np.random.default_rng(0)
    n, me = 40, 12
    M = rng.standard_normal((n, n))
    P = M @ M.T + np.eye(n)
    q = rng.standard_normal(n)
    A = rng.standard_normal((me, n))
    b = rng.standard_normal(me)
    K = np.block([[P, A.T], [A, np.zeros((me, me))]])
    x_star = np.linalg.solve(K, np.concatenate([-q, b]))[:n]
    x, y, info = op.osqp_solve(sp.csc_matrix(P), q, sp.csc_matrix(A), b, b, eps_abs=1e-9, eps_rel=1e-9)
    assert info["status"] == "solved"
    np.testing.assert_allclose(x, x_star, rtol=0, atol=1e-6)


def test_box_qp_satisfies_kkt():
    import osqp_port as op

    rng = np.random.default_rng(1)
    n = 30
    M = rng.standard_normal((n, n))
    P = M @ M.T + 0.1 * np.eye(n)
    q = 3 * rng.standard_normal(n)
    lo, hi = -0.3 * np.ones(n), 0.5 * np.ones(n)
    x, y, info = op.osqp_solve(sp.csc_matrix(P), q, sp.identity(n, format="csc"), lo, hi, eps_abs=1e-8, eps_rel=1e-8)
    assert info["status"] == "solved" and (x >= lo - 1e-6).all() and (x <= hi + 1e-6).all()
    g = P @ x + q  # gradient: zero on free variables, pushes outwards on active bounds
    free = (x > lo + 1e-5) & (x < hi - 1e-5)
    assert free.any() and (~free).any()
    assert np.abs(g[free]).max() < 1e-4 and (g[x <= lo + 1e-5] >= -1e-4).all() and (g[x >= hi - 1e-5] <= 1e-4).all()


def test_first_order_solution_at_cvxpy_tolerances_vs_high_accuracy_oracle(capsys):
    """Default-size DD-MPC problem (n=6, L=35, N=350) on the golden drone data: both QPs of a step solved at
    eps_abs = eps_rel = 1e-5 (what cvxpy asks of OSQP) terminate as `solved`, yet the optimal input lies ~1e-3..1e-2 from
    the KKT-certified optimum - two orders above the north star's 1e-4, which can therefore only be asked of a
    high-accuracy solution.  The GPU solver is held to 1e-5 against that optimum (tests/test_ddmpc_parity_gpu.py)."""
    import ddmpc_oracle as mo
    import osqp_port as op

    g = np.load(os.path.join(GOLDEN, "ddmpc_default.npz"))
    ctrl = op.OSQPStyleDDMPC(mo.default_params(), g["u"], g["y"], g["y_r"])
    worst0 = worst = 0.0
    for t in range(2):
        ctrl.update_and_solve_data_driven_mpc()
        assert ctrl.info["mpc"]["status"] == "solved" and ctrl.info["alpha_sr"]["status"] == "solved"
        assert ctrl.info["mpc"]["iters"] <= 2000
        worst0 = max(worst0, float(np.abs(ctrl.optimal_u[0] - g["u_opt"][t][0]).max()))
        worst = max(worst, float(np.abs(ctrl.optimal_u - g["u_opt"][t]).max()))
        ctrl.store_input_output_measurement(g["u_opt"][t][0], g["y_next"][t])
    with capsys.disabled():
        print(f"\n[osqp-port] |u*_0 - oracle| = {worst0:.2e}, max over the horizon = {worst:.2e} at eps 1e-5")
    assert 1e-5 < worst0 < 5e-2  # close to the optimum, but nowhere near 1e-4: the tolerance of the solver class decides


def test_tight_tolerances_converge_to_the_oracle():
    """With eps = 1e-9 the same port reaches the oracle's optimum on the reference's test-sized problem
    (tests/data_driven_mpc_tests/config/test_controller_params.yaml: n=2, L=5)."""
    import ddmpc_oracle as mo
    import osqp_port as op
    from test_ddmpc_parity_gpu import _small_params, _synthetic_windows

    prm = _small_params()
    prm.alpha_reg_type = 2
    u, y = _synthetic_windows(prm, 5, 1)
    y_r = np.array([0.3, -0.2, 0.5])
    orc = mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u[0], y[0], y_r))
    ctrl = op.OSQPStyleDDMPC(prm, u[0], y[0], y_r, eps_abs=1e-10, eps_rel=1e-10, max_iter=40000)
    ctrl.update_and_solve_data_driven_mpc()
    assert ctrl.info["mpc"]["status"] == "solved"
    np.testing.assert_allclose(ctrl.optimal_u, orc.optimal_u, rtol=0, atol=2e-5)
