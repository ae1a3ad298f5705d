"""GPU parity of the batched condensed DD-MPC solver against the literal-QP interior-point oracle."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north star: optimal inputs within 1e-4 of the reference solution (problem conditioning ~1e11 limits fp64 agreement to ~1e-6)


def _small_params():
    import ddmpc_oracle as mo

    # tests/data_driven_mpc_tests/config/test_controller_params.yaml of the reference: N=40, n=2, L=5
    prm = mo.default_params()
    prm.n, prm.L, prm.N = 2, 5, 80
    return prm


class _Plant:
    """Stable random linear system x+ = A x + B u, y = C x + noise (one per controller)."""

    def __init__(self, rng, m, p):
        self.A = np.diag(rng.uniform(0.6, 0.95, 6)) + 0.02 * rng.standard_normal((6, 6))
        self.B = rng.standard_normal((6, m)) * 0.3
        self.C = rng.standard_normal((p, 6)) * 0.5
        self.x = np.zeros(6)
        self.rng, self.p = rng, p

    def step(self, u):
        self.x = self.A @ self.x + self.B @ u
        return self.C @ self.x + 1e-4 * self.rng.standard_normal(self.p)


def _synthetic_windows(prm, seed, B, return_plants=False):
    """Cheap persistently exciting data: stable random linear systems driven by bounded noise."""
    rng = np.random.default_rng(seed)
    us, ys, plants = [], [], []
    for _ in range(B):
        plant = _Plant(rng, prm.m, prm.p)
        u = np.stack([rng.uniform(0.15, 0.4, prm.N), rng.uniform(-0.5, 0.5, prm.N), rng.uniform(-0.5, 0.5, prm.N)], 1)
        y = np.stack([plant.step(u[k]) for k in range(prm.N)])
        us.append(u)
        ys.append(y)
        plants.append(plant)
    if return_plants:
        return np.stack(us), np.stack(ys), plants
    return np.stack(us), np.stack(ys)


def _batched(prm, u, y, y_r, **kw):
    import ddmpc_oracle as mo
    from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import (
        BatchedNonlinearDataDrivenMPC,
    )

    kwargs = mo.ctor_kwargs(prm, u, y, y_r[0])
    kwargs["y_r"] = y_r
    return BatchedNonlinearDataDrivenMPC(**{**kwargs, "device": "cuda", **kw})


@pytest.mark.parametrize("alpha_reg_type", [0, 2])
def test_small_batch_closed_loop_matches_oracle(build_lib, alpha_reg_type):
    import ddmpc_oracle as mo

    prm = _small_params()
    prm.alpha_reg_type = alpha_reg_type
    B, T = 6, 4
    u, y, plants = _synthetic_windows(prm, 3, B, return_plants=True)
    y_r = np.tile(np.array([[0.3, -0.2, 0.5]]), (B, 1)) + 0.1 * np.arange(B)[:, None]
    gpu = _batched(prm, u, y, y_r, solve_on_init=False)
    orcs = [mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u[b], y[b], y_r[b]), solve_on_init=False) for b in range(B)]
    for t in range(T):
        gpu.update_and_solve_data_driven_mpc()
        assert (gpu.status.cpu().numpy() == 0).all(), gpu.status
        u0 = gpu.get_optimal_control_input_at_step(0).cpu().numpy()
        upred = gpu.optimal_u.cpu().numpy()
        y_new = np.stack([plants[b].step(u0[b]) for b in range(B)])  # closed loop on the true plants
        for b, orc in enumerate(orcs):
            orc.update_and_solve_data_driven_mpc()
            np.testing.assert_allclose(upred[b], orc.optimal_u, rtol=0, atol=TOL)
            np.testing.assert_allclose(gpu.us[b].cpu().numpy(), orc.us_prev, rtol=0, atol=TOL)
            np.testing.assert_allclose(gpu.ys[b].cpu().numpy(), orc.ys_prev, rtol=0, atol=10 * TOL)
            orc.store_input_output_measurement(u0[b], y_new[b])
        gpu.store_input_output_measurement(torch.from_numpy(u0).cuda(), torch.from_numpy(y_new).cuda())
        for b, orc in enumerate(orcs):  # rolling windows stay identical
            assert np.array_equal(gpu.u[b].cpu().numpy(), orc.u) and np.array_equal(gpu.y[b].cpu().numpy(), orc.y)


def test_default_size_matches_oracle_and_golden(build_lib):
    """n=6, L=35, N=350 (693-variable QP) on PID-stabilised drone data (tests/golden/ddmpc_default.npz)."""
    import os

    import ddmpc_oracle as mo

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddmpc_default.npz"))
    prm = mo.default_params()
    u, y, y_r = g["u"], g["y"], g["y_r"]
    gpu = _batched(prm, u[None], y[None], y_r[None], solve_on_init=False)
    for t in range(g["u_opt"].shape[0]):
        gpu.update_and_solve_data_driven_mpc()
        assert int(gpu.status[0]) == 0
        np.testing.assert_allclose(gpu.optimal_u[0].cpu().numpy(), g["u_opt"][t], rtol=0, atol=TOL)
        np.testing.assert_allclose(gpu.us[0].cpu().numpy(), g["us"][t], rtol=0, atol=TOL)
        gpu.store_input_output_measurement(g["u_opt"][t][0][None], g["y_next"][t][None])
    assert int(gpu.pivoting_passes.max()) <= 20 and int(gpu.ipm_iterations.max()) == 0


def test_single_controller_numpy_api(build_lib):
    import ddmpc_oracle as mo
    from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import (
        NonlinearDataDrivenMPCController,
    )

    prm = _small_params()
    u, y = _synthetic_windows(prm, 11, 1)
    y_r = np.array([0.2, 0.1, 0.4])
    ctrl = NonlinearDataDrivenMPCController(**mo.ctor_kwargs(prm, u[0], y[0], y_r))
    orc = mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u[0], y[0], y_r))
    u0 = ctrl.get_optimal_control_input_at_step(n_step=0)
    assert isinstance(u0, np.ndarray) and u0.shape == (3,)
    np.testing.assert_allclose(u0, orc.get_optimal_control_input_at_step(0), rtol=0, atol=TOL)
    assert ctrl.get_du_value_at_step(0) is None and ctrl.n == prm.n and ctrl.y_r.shape == (3, 1)
    ctrl.set_output_setpoint(np.array([[0.0], [0.0], [0.1]]))
    orc.set_output_setpoint(np.array([[0.0], [0.0], [0.1]]))
    ctrl.store_input_output_measurement(u0, y[0, -1])
    orc.store_input_output_measurement(u0, y[0, -1])
    ctrl.update_and_solve_data_driven_mpc()
    orc.update_and_solve_data_driven_mpc()
    np.testing.assert_allclose(ctrl.get_optimal_control_input_at_step(0), orc.get_optimal_control_input_at_step(0), rtol=0, atol=TOL)
    with pytest.raises(ValueError):
        ctrl.get_optimal_control_input_at_step(n_step=1)


def test_unsupported_modes_fail_loudly(build_lib):
    import ddmpc_oracle as mo
    from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import (
        NonlinearDataDrivenMPCController,
    )

    prm = _small_params()
    u, y = _synthetic_windows(prm, 1, 1)
    kw = mo.ctor_kwargs(prm, u[0], y[0], np.zeros(3))
    with pytest.raises(NotImplementedError):
        NonlinearDataDrivenMPCController(**{**kw, "ext_out_incr_in": True})
    Q_full = kw["Q"].copy()
    Q_full[0, 1] = Q_full[1, 0] = 0.5  # not kron(I, diag(w))
    with pytest.raises(NotImplementedError):
        NonlinearDataDrivenMPCController(**{**kw, "Q": Q_full})
    with pytest.raises(ValueError):
        NonlinearDataDrivenMPCController(**{**kw, "u": u[0][:30], "y": y[0][:30]})


def test_hard_cases_fall_back_to_interior_point_and_match_oracle(build_lib):
    """Saved default-size windows far from the data's operating point (2-39 active bounds): every controller
    must converge (status 0) and match the literal-QP oracle solutions of the fixture, also when the pivoting
    is cut short and the interior-point fallback + polish has to finish the job."""
    import os

    import ddmpc_oracle as mo

    h = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddmpc_hard_cases.npz"))
    B = h["u"].shape[0]
    prm = mo.default_params()
    gpu = _batched(prm, h["u"], h["y"], np.tile(h["y_r"], (B, 1)), solve_on_init=False, warm_start=False)
    gpu.update_and_solve_data_driven_mpc()
    assert (gpu.status.cpu().numpy() == 0).all(), gpu.status
    np.testing.assert_allclose(gpu.optimal_u.cpu().numpy(), h["u_opt"], rtol=0, atol=TOL)
    # force the interior-point fallback (pass cap 1): same optimum through the other path
    forced = _batched(prm, h["u"], h["y"], np.tile(h["y_r"], (B, 1)), solve_on_init=False, warm_start=False, max_iter=1)
    forced.update_and_solve_data_driven_mpc()
    assert (forced.status.cpu().numpy() == 0).all() and int((forced.ipm_iterations > 0).sum()) >= B // 2
    np.testing.assert_allclose(forced.optimal_u.cpu().numpy(), h["u_opt"], rtol=0, atol=TOL)
    np.testing.assert_allclose(gpu.optimal_u.cpu().numpy(), h["u_opt"], rtol=0, atol=TOL)
    np.testing.assert_allclose(gpu.us.cpu().numpy(), h["us"], rtol=0, atol=TOL)


def test_warm_start_does_not_change_the_solution(build_lib):
    """The carried active set only seeds the pivoting: warm- and cold-started controllers must return the
    same optimum at every closed-loop step."""
    prm = _small_params()
    prm.alpha_reg_type = 0
    B, T = 5, 6
    u, y, plants = _synthetic_windows(prm, 17, B, return_plants=True)
    y_r = np.tile(np.array([[0.4, -0.3, 0.6]]), (B, 1))
    warm = _batched(prm, u, y, y_r, solve_on_init=False, warm_start=True)
    cold = _batched(prm, u, y, y_r, solve_on_init=False, warm_start=False)
    for t in range(T):
        warm.update_and_solve_data_driven_mpc()
        cold.update_and_solve_data_driven_mpc()
        assert (warm.status.cpu().numpy() == 0).all() and (cold.status.cpu().numpy() == 0).all()
        np.testing.assert_allclose(warm.optimal_u.cpu().numpy(), cold.optimal_u.cpu().numpy(), rtol=0, atol=1e-9)
        u0 = warm.get_optimal_control_input_at_step(0).cpu().numpy()
        y_new = np.stack([plants[b].step(u0[b]) for b in range(B)])
        for c in (warm, cold):
            c.store_input_output_measurement(torch.from_numpy(u0).cuda(), torch.from_numpy(y_new).cuda())
    assert int(warm.pivoting_passes.max()) <= int(cold.pivoting_passes.max()) + 2


@pytest.mark.parametrize("n,L,N", [(2, 6, 80), (2, 7, 90), (3, 6, 100), (3, 9, 130), (4, 10, 160)])
def test_other_problem_shapes_match_oracle(build_lib, n, L, N):
    """Shapes with different paddings of the row classes (R2 / R3 not multiples of 8, C a multiple of 8, odd l, a
    box-QP factor larger than the dense Schur copy): the grid search sweeps n and L."""
    import ddmpc_oracle as mo

    prm = mo.default_params()
    prm.n, prm.L, prm.N, prm.alpha_reg_type = n, L, N, 0
    u, y = _synthetic_windows(prm, 3, 2)
    y_r = np.tile(np.array([[0.3, -0.2, 0.5]]), (2, 1))
    gpu = _batched(prm, u, y, y_r, solve_on_init=False)
    orcs = [mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u[b], y[b], y_r[b]), solve_on_init=False) for b in range(2)]
    for t in range(2):
        gpu.update_and_solve_data_driven_mpc()
        assert (gpu.status.cpu().numpy() == 0).all()
        for b, orc in enumerate(orcs):
            orc.update_and_solve_data_driven_mpc()
            np.testing.assert_allclose(gpu.optimal_u[b].cpu().numpy(), orc.optimal_u, rtol=0, atol=TOL)
            orc.store_input_output_measurement(orc.optimal_u[0], y[b, -1])
        gpu.store_input_output_measurement(torch.from_numpy(np.stack([o.u[-1] for o in orcs])).cuda(), torch.from_numpy(y[:, -1].copy()).cuda())


def test_full_batch_4096_default_size(build_lib):
    """BASELINE configs[3] scale: 4096 default-size controllers in one launch.  Size-independent properties - every
    controller converges, the result of a controller does not depend on its place in the batch (a permuted batch
    returns the permuted result bit for bit: persistent CTAs pull controllers from a work counter, so the CTA,
    workspace slot and order that serve a controller differ between the two launches) - plus the literal-QP
    oracle on a sample."""
    import os

    import ddmpc_oracle as mo

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddmpc_default.npz"))
    prm = mo.default_params()
    B = 4096
    rng = np.random.default_rng(7)
    u = np.repeat(g["u"][None], B, 0) + 1e-3 * rng.standard_normal((B,) + g["u"].shape)
    y = np.repeat(g["y"][None], B, 0) + 1e-4 * rng.standard_normal((B,) + g["y"].shape)
    y_r = np.tile(g["y_r"][None], (B, 1)) + 0.2 * rng.uniform(-1, 1, (B, 3))
    gpu = _batched(prm, u, y, y_r, solve_on_init=False)
    gpu.update_and_solve_data_driven_mpc()
    assert (gpu.status.cpu().numpy() == 0).all(), int((gpu.status != 0).sum())
    uo = gpu.optimal_u.cpu().numpy()
    lo, hi = np.asarray(prm.U)[:, 0], np.asarray(prm.U)[:, 1]
    assert np.isfinite(uo).all() and (uo >= lo - 1e-9).all() and (uo <= hi + 1e-9).all()
    perm = rng.permutation(B)
    gpu2 = _batched(prm, u[perm], y[perm], y_r[perm], solve_on_init=False)
    gpu2.update_and_solve_data_driven_mpc()
    assert np.array_equal(gpu2.optimal_u.cpu().numpy(), uo[perm])
    assert np.array_equal(gpu2.us.cpu().numpy(), gpu.us.cpu().numpy()[perm])
    for b in (0, 1733, 4095):
        orc = mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u[b], y[b], y_r[b]))
        np.testing.assert_allclose(uo[b], orc.optimal_u, rtol=0, atol=TOL)


@pytest.mark.parametrize("shape", ["small", "default", "L30"])
def test_gram_cache_follows_the_window_like_a_regeneration(build_lib, shape):
    """The per-controller Gram cache (rank-2 correction per one-sample shift) against the regeneration of G from the
    window in a closed loop: same optimal inputs to 1e-6 (the two differ by rounding of order ulp(G_ij) only), through
    every cache transition - first solve (fill), one shift (update), refresh after `gram_refresh` updates, a second
    solve without a shift (reuse), two shifts between solves (regenerate), direct window write + invalidate."""
    import os

    import ddmpc_oracle as mo

    if shape == "small":
        prm, B = _small_params(), 6
        u, y, plants = _synthetic_windows(prm, 3, B, return_plants=True)
        y_r = np.tile(np.array([[0.2, -0.1, 0.1]]), (B, 1))
    else:
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddmpc_default.npz"))
        prm, B = mo.default_params(), 5
        if shape == "L30":  # second shape bucket of the reference's grid (dd_mpc_grid_search_params.yaml:93-100)
            prm.L = 30
        rng = np.random.default_rng(2)
        u = np.repeat(g["u"][None], B, 0) + 1e-3 * rng.standard_normal((B,) + g["u"].shape)
        y = np.repeat(g["y"][None], B, 0) + 1e-4 * rng.standard_normal((B,) + g["y"].shape)
        y_r = np.tile(g["y_r"][None], (B, 1))
        plants = None
    a = _batched(prm, u, y, y_r, solve_on_init=False, gram_cache=True, gram_refresh=4)
    b = _batched(prm, u, y, y_r, solve_on_init=False, gram_cache=False)
    assert a.gram_cache_ptr is not None and b.gram_cache_ptr is None
    rng = np.random.default_rng(9)
    worst = 0.0

    def solve_both():
        nonlocal worst
        a.update_and_solve_data_driven_mpc()
        b.update_and_solve_data_driven_mpc()
        assert (a.status.cpu().numpy() == 0).all() and (b.status.cpu().numpy() == 0).all()
        ua, ub = a.optimal_u.cpu().numpy(), b.optimal_u.cpu().numpy()
        worst = max(worst, float(np.abs(ua - ub).max()))
        np.testing.assert_allclose(ua, ub, rtol=0, atol=1e-6)

    def store_both():
        u0 = b.get_optimal_control_input_at_step(0).cpu().numpy()
        if plants is not None:
            yk = np.stack([plants[i].step(u0[i]) for i in range(B)])
        else:
            yk = b.y[:, -1].cpu().numpy() + 1e-3 * rng.standard_normal((B, prm.p))
        for c in (a, b):
            c.store_input_output_measurement(torch.from_numpy(u0).cuda(), torch.from_numpy(yk).cuda())

    solve_both()  # fill
    for _ in range(11):  # update x4, regenerate, update ...
        store_both()
        solve_both()
    solve_both()  # reuse (no shift since the last solve)
    store_both()
    store_both()
    solve_both()  # two shifts: regenerate
    assert torch.equal(a.u, b.u) and torch.equal(a.y, b.y)
    for c in (a, b):
        c.y[:, -1] += 1e-3
    a.invalidate_gram_cache()
    solve_both()
    print(f"gram cache vs regeneration ({shape}): max |du| = {worst:.3e}")
    if shape != "small":  # (small shapes have no room for the stages: the kernel regenerates G and the runs are identical)
        assert worst > 0.0, "the cached path was not exercised"


@pytest.mark.parametrize("n,L", [(4, 30), (4, 35), (6, 30)])
def test_reference_grid_shapes_match_oracle(build_lib, n, L):
    """The other three QP shapes of the reference's parameter grid (configs/data_driven_mpc/dd_mpc_grid_search_params.yaml:
    93-100: n in {4, 6}, L in {30, 35}; N = 350) on the PID-stabilised drone data, three closed-loop steps (the second and
    third on the Gram-cache path) against the literal-QP oracle."""
    import os

    import ddmpc_oracle as mo

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddmpc_default.npz"))
    prm = mo.default_params()
    prm.n, prm.L = n, L
    u, y, y_r = g["u"], g["y"], g["y_r"]
    gpu = _batched(prm, u[None], y[None], y_r[None], solve_on_init=False)
    orc = mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u, y, y_r), solve_on_init=False)
    rng = np.random.default_rng(4)
    for t in range(3):
        gpu.update_and_solve_data_driven_mpc()
        orc.update_and_solve_data_driven_mpc()
        assert int(gpu.status[0]) == 0
        np.testing.assert_allclose(gpu.optimal_u[0].cpu().numpy(), orc.optimal_u, rtol=0, atol=TOL)
        y_new = orc.y[-1] + 1e-3 * rng.standard_normal(3)
        u0 = orc.optimal_u[0].copy()
        orc.store_input_output_measurement(u0, y_new)
        gpu.store_input_output_measurement(u0[None], y_new[None])


def test_gram_cache_drift_over_a_refresh_period(build_lib):
    """140 consecutive one-sample shifts with the default refresh period (128 rank-2 corrections between
    regenerations): the optimal inputs of the cached path stay within 1e-6 of the regenerating path at every step,
    i.e. the rounding of the corrections does not accumulate into the solution."""
    import os

    import ddmpc_oracle as mo

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddmpc_default.npz"))
    prm, B = mo.default_params(), 4
    rng = np.random.default_rng(12)
    u = np.repeat(g["u"][None], B, 0) + 1e-3 * rng.standard_normal((B,) + g["u"].shape)
    y = np.repeat(g["y"][None], B, 0) + 1e-4 * rng.standard_normal((B,) + g["y"].shape)
    y_r = np.tile(g["y_r"][None], (B, 1))
    a = _batched(prm, u, y, y_r, solve_on_init=False, gram_cache=True)
    b = _batched(prm, u, y, y_r, solve_on_init=False, gram_cache=False)
    worst = []
    for t in range(140):
        a.update_and_solve_data_driven_mpc()
        b.update_and_solve_data_driven_mpc()
        worst.append(float((a.optimal_u - b.optimal_u).abs().max()))
        # a smooth excursion of the outputs and the applied optimal input: the window keeps changing
        u0 = b.get_optimal_control_input_at_step(0).clone()
        yk = b.y[:, -1] + torch.from_numpy(2e-3 * np.sin(0.1 * t) + 1e-4 * rng.standard_normal((B, 3))).cuda()
        for c in (a, b):
            c.store_input_output_measurement(u0, yk)
    assert int((a.status != 0).sum()) == 0 and int((b.status != 0).sum()) == 0
    print(f"gram cache drift: max |du| first 10 steps {max(worst[:10]):.2e}, steps 118-127 {max(worst[118:128]):.2e}, "
          f"after the regeneration {max(worst[129:]):.2e}")
    assert max(worst) <= 1e-6


def test_skipped_controllers_are_left_untouched(build_lib):
    """`update_and_solve_data_driven_mpc(skip=mask)` (have_prev < 0 in the C ABI): the masked controllers keep outputs,
    status, active set and Gram cache; the others get exactly the result of an unmasked solve; a controller that was
    skipped can be solved again later and then agrees with one that never was."""
    prm, B = _small_params(), 6
    prm.n, prm.L, prm.N = 4, 10, 160  # large enough for the Gram-cache path
    u, y, plants = _synthetic_windows(prm, 5, B, return_plants=True)
    y_r = np.tile(np.array([[0.2, -0.1, 0.1]]), (B, 1))
    a = _batched(prm, u, y, y_r, solve_on_init=False)
    b = _batched(prm, u, y, y_r, solve_on_init=False)
    skip = torch.tensor([False, True, False, True, True, False], device="cuda")
    for c in (a, b):
        c.update_and_solve_data_driven_mpc()
    for step in range(3):
        u0 = b.get_optimal_control_input_at_step(0).clone()
        yk = torch.from_numpy(np.stack([plants[i].step(u0[i].cpu().numpy()) for i in range(B)])).cuda()
        for c in (a, b):
            c.store_input_output_measurement(u0, yk)
        before = a.optimal_u.clone()
        a.update_and_solve_data_driven_mpc(skip=skip if step < 2 else None)
        b.update_and_solve_data_driven_mpc()
        if step < 2:
            assert torch.equal(a.optimal_u[skip], before[skip])
            assert torch.equal(a.optimal_u[~skip], b.optimal_u[~skip])
        else:  # solved again after two skipped steps (Gram regenerated: the window moved by more than one sample)
            np.testing.assert_allclose(a.optimal_u.cpu().numpy(), b.optimal_u.cpu().numpy(), rtol=0, atol=1e-6)
        assert int((a.status != 0).sum()) == 0


def test_previous_equilibrium_anchor_matches_oracle_small(build_lib):
    """alpha_sr_anchor="previous_equilibrium" (the reading of alpha_reg_type APPROXIMATED recalled in SURVEY 8 a15):
    alpha_sr = 0 on the first solve, then the alpha_sr of the previous solve's optimal (u_s, y_s).  Closed loop of 5
    steps on 6 controllers against the literal-QP oracle in the same mode; the two anchors really differ."""
    import ddmpc_oracle as mo

    prm = _small_params()
    prm.alpha_reg_type = 0
    B, T = 6, 5
    u, y, plants = _synthetic_windows(prm, 3, B, return_plants=True)
    y_r = np.tile(np.array([[0.3, -0.2, 0.5]]), (B, 1)) + 0.1 * np.arange(B)[:, None]
    gpu = _batched(prm, u, y, y_r, solve_on_init=False, alpha_sr_anchor="previous_equilibrium")
    ref = _batched(prm, u, y, y_r, solve_on_init=False)
    orcs = [mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u[b], y[b], y_r[b]), solve_on_init=False,
                                    alpha_sr_anchor="previous_equilibrium") for b in range(B)]  # fmt: skip
    differs = 0.0
    for t in range(T):
        gpu.update_and_solve_data_driven_mpc()
        assert (gpu.status.cpu().numpy() == 0).all(), gpu.status
        u0 = gpu.get_optimal_control_input_at_step(0).cpu().numpy()
        upred = gpu.optimal_u.cpu().numpy()
        y_new = np.stack([plants[b].step(u0[b]) for b in range(B)])
        for b, orc in enumerate(orcs):
            orc.update_and_solve_data_driven_mpc()
            np.testing.assert_allclose(upred[b], orc.optimal_u, rtol=0, atol=TOL)
            np.testing.assert_allclose(gpu.us[b].cpu().numpy(), orc.us_prev, rtol=0, atol=TOL)
            orc.store_input_output_measurement(u0[b], y_new[b])
        if t == 0:
            ref.update_and_solve_data_driven_mpc()
            differs = float(np.abs(ref.optimal_u.cpu().numpy() - upred).max())
        gpu.store_input_output_measurement(torch.from_numpy(u0).cuda(), torch.from_numpy(y_new).cuda())
    assert differs > 1e-4  # anchor 1's first solve (alpha_sr = 0) is not the optimal-reachable-equilibrium solution


def test_previous_equilibrium_anchor_matches_oracle_default_size(build_lib):
    """The same at n=6, L=35, N=350 on the PID-stabilised drone data of the golden fixture (2 closed-loop steps)."""
    import os

    import ddmpc_oracle as mo

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddmpc_default.npz"))
    prm = mo.default_params()
    u, y, y_r = g["u"], g["y"], g["y_r"]
    gpu = _batched(prm, u[None], y[None], y_r[None], solve_on_init=False, alpha_sr_anchor="previous_equilibrium")
    orc = mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u, y, y_r), solve_on_init=False, alpha_sr_anchor="previous_equilibrium")
    for t in range(2):
        gpu.update_and_solve_data_driven_mpc()
        orc.update_and_solve_data_driven_mpc()
        assert int(gpu.status[0]) == 0
        np.testing.assert_allclose(gpu.optimal_u[0].cpu().numpy(), orc.optimal_u, rtol=0, atol=TOL)
        np.testing.assert_allclose(gpu.us[0].cpu().numpy(), orc.us_prev, rtol=0, atol=TOL)
        u0 = orc.optimal_u[0]
        gpu.store_input_output_measurement(u0[None], g["y_next"][t][None])
        orc.store_input_output_measurement(u0, g["y_next"][t])


# ---- the optional modes of SURVEY 8 a16: alpha_reg_type PREVIOUS, update_cost_threshold (dual recovery kernel) ----------


def _closed_loop_against_oracle(prm, B, T, seed, gpu_kw, orc_kw, check_alpha=False, tol=TOL):
    import ddmpc_oracle as mo

    u, y, plants = _synthetic_windows(prm, seed, B, return_plants=True)
    y_r = np.tile(np.array([[0.3, -0.2, 0.5]]), (B, 1)) + 0.1 * np.arange(B)[:, None]
    gpu = _batched(prm, u, y, y_r, solve_on_init=False, **gpu_kw)
    orcs = [mo.NonlinearDDMPCOracle(**{**mo.ctor_kwargs(prm, u[b], y[b], y_r[b]), **orc_kw}, solve_on_init=False) for b in range(B)]
    for t in range(T):
        gpu.update_and_solve_data_driven_mpc()
        assert (gpu.status.cpu().numpy() == 0).all(), gpu.status
        u0 = gpu.get_optimal_control_input_at_step(0).cpu().numpy()
        upred = gpu.optimal_u.cpu().numpy()
        y_new = np.stack([plants[b].step(u0[b]) for b in range(B)])
        for b, orc in enumerate(orcs):
            orc.update_and_solve_data_driven_mpc()
            np.testing.assert_allclose(upred[b], orc.optimal_u, rtol=0, atol=tol, err_msg=f"step {t} controller {b}")
            np.testing.assert_allclose(gpu.us[b].cpu().numpy(), orc.us_prev, rtol=0, atol=tol)
            if check_alpha:  # the literal QP is strictly convex in alpha: alpha* is unique
                np.testing.assert_allclose(gpu.alpha_prev[b].cpu().numpy(), orc.alpha_prev, rtol=0, atol=tol)
            orc.store_input_output_measurement(u0[b], y_new[b])
        yield gpu, orcs, t
        gpu.store_input_output_measurement(torch.from_numpy(u0).cuda(), torch.from_numpy(y_new).cuda())


def test_alpha_reg_type_previous_matches_oracle_small(build_lib):
    """alpha_reg_type = 1 (configs/data_driven_mpc/dd_mpc_controller_params.yaml:60-66): alpha_sr = the previous optimal
    alpha.  The condensed solver never forms alpha; `ddqc_ddmpc_recover` recovers alpha* = alpha_sr + H'z after every
    solve.  6 controllers x 6 closed-loop steps against the literal-QP oracle: optimal inputs AND the recovered alpha."""
    prm = _small_params()
    prm.alpha_reg_type = 1
    zero = None
    for gpu, orcs, t in _closed_loop_against_oracle(prm, 6, 6, 3, {}, {}, check_alpha=True):
        if t == 1:  # from the second solve on the answer differs from alpha_reg_type ZERO
            prm0 = _small_params()
            prm0.alpha_reg_type = 2
            zero = _batched(prm0, gpu.u.cpu().numpy(), gpu.y.cpu().numpy(), gpu.y_r.cpu().numpy())
            assert float((zero.optimal_u - gpu.optimal_u).abs().max()) > 1e-4
    assert zero is not None


def test_alpha_reg_type_previous_matches_oracle_default_size(build_lib):
    """The same at n=6, L=35, N=350 (alpha in R^309) on the PID-stabilised drone data of the golden fixture, with and
    without the Gram cache (3 closed-loop steps)."""
    import os

    import ddmpc_oracle as mo

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddmpc_default.npz"))
    prm = mo.default_params()
    prm.alpha_reg_type = 1
    u, y, y_r = g["u"], g["y"], g["y_r"]
    gpus = [_batched(prm, u[None], y[None], y_r[None], solve_on_init=False, gram_cache=gc) for gc in (True, False)]
    orc = mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u, y, y_r), solve_on_init=False)
    for t in range(3):
        orc.update_and_solve_data_driven_mpc()
        for gpu in gpus:
            gpu.update_and_solve_data_driven_mpc()
            assert int(gpu.status[0]) == 0
            # 5e-5 instead of TOL: at step 2 one predicted input sits 1.3e-5 from its bound in the ORACLE's interior-point
            # iterate (weakly active, mu = 6e-11) while the pivoting solver puts it on the bound; the float64 numpy mirror
            # of the condensed algorithm gives the GPU's value to 1e-8 (measured, tools note in DESIGN section 4d)
            np.testing.assert_allclose(gpu.optimal_u[0].cpu().numpy(), orc.optimal_u, rtol=0, atol=5e-5)
            np.testing.assert_allclose(gpu.alpha_prev[0].cpu().numpy(), orc.alpha_prev, rtol=0, atol=5e-5)
            gpu.store_input_output_measurement(orc.optimal_u[0][None], g["y_next"][t][None])
        orc.store_input_output_measurement(orc.optimal_u[0], g["y_next"][t])


@pytest.mark.parametrize("alpha_reg_type", [0, 1, 2])
def test_update_cost_threshold_matches_oracle(build_lib, alpha_reg_type):
    """update_cost_threshold (dd_mpc_controller_params.yaml:85-90): the Hankel data of a controller stops following the
    measurements while its tracking cost is below the threshold (the initial condition keeps following them).  The
    tracking cost comes from the recovered dual; data-update decisions, costs and optimal inputs against the oracle."""
    prm = _small_params()
    prm.alpha_reg_type = alpha_reg_type
    thr = {0: 9.0, 1: 20.0, 2: 55.0}[alpha_reg_type]
    froze = updated = 0
    for gpu, orcs, t in _closed_loop_against_oracle(prm, 6, 8, 3, dict(update_cost_threshold=thr), dict(update_cost_threshold=thr),
                                                    check_alpha=alpha_reg_type == 1):
        cost = gpu.tracking_cost.cpu().numpy()
        upd = gpu.data_updated.cpu().numpy()
        for b, orc in enumerate(orcs):
            assert abs(cost[b] - orc.tracking_cost) <= 1e-6 * max(1.0, orc.tracking_cost), (t, b, cost[b], orc.tracking_cost)
            assert bool(upd[b]) == bool(orc.data_updated), (t, b)
            assert np.array_equal(gpu.u_data[b].cpu().numpy(), orc.u_data) and np.array_equal(gpu.y_data[b].cpu().numpy(), orc.y_data)
            froze += not orc.data_updated
            updated += bool(orc.data_updated) and t > 0
    assert froze > 0 and updated > 0, (froze, updated)  # the test really exercises both branches


def test_update_cost_threshold_default_size(build_lib):
    """Default sizes, alpha_reg_type APPROXIMATED (the shipped setting) with a threshold that freezes the data after the
    first solve: the frozen-data solve (snapshot + fresh initial condition) against the oracle."""
    import os

    import ddmpc_oracle as mo

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddmpc_default.npz"))
    prm = mo.default_params()
    u, y, y_r = g["u"], g["y"], g["y_r"]
    kw = dict(update_cost_threshold=1e9)
    gpu = _batched(prm, u[None], y[None], y_r[None], solve_on_init=False, **kw)
    orc = mo.NonlinearDDMPCOracle(**{**mo.ctor_kwargs(prm, u, y, y_r), **kw}, solve_on_init=False)
    for t in range(2):
        orc.update_and_solve_data_driven_mpc()
        gpu.update_and_solve_data_driven_mpc()
        assert int(gpu.status[0]) == 0 and bool(gpu.data_updated[0]) == orc.data_updated == (t == 0)
        np.testing.assert_allclose(gpu.optimal_u[0].cpu().numpy(), orc.optimal_u, rtol=0, atol=TOL)
        assert abs(float(gpu.tracking_cost[0]) - orc.tracking_cost) <= 1e-6 * orc.tracking_cost
        gpu.store_input_output_measurement(orc.optimal_u[0][None], g["y_next"][t][None])
        orc.store_input_output_measurement(orc.optimal_u[0], g["y_next"][t])


def test_single_controller_numpy_api_optional_modes(build_lib):
    """The numpy-in / numpy-out drop-in (`NonlinearDataDrivenMPCController`, the class the reference constructs at
    controller_creation.py:91-112) with `alpha_reg_type=PREVIOUS` and `update_cost_threshold` passed as the reference passes
    them: three closed-loop steps against the oracle."""
    import ddmpc_oracle as mo
    from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import (
        AlphaRegType,
        NonlinearDataDrivenMPCController,
    )

    prm = _small_params()
    prm.alpha_reg_type = 1
    u, y, plants = _synthetic_windows(prm, 21, 1, return_plants=True)
    y_r = np.array([0.25, -0.1, 0.45])
    kw = {**mo.ctor_kwargs(prm, u[0], y[0], y_r), "update_cost_threshold": 25.0}
    ctrl = NonlinearDataDrivenMPCController(**{**kw, "alpha_reg_type": AlphaRegType.PREVIOUS})
    orc = mo.NonlinearDDMPCOracle(**kw)
    for t in range(3):
        u0 = ctrl.get_optimal_control_input_at_step(n_step=0)
        np.testing.assert_allclose(u0, orc.get_optimal_control_input_at_step(0), rtol=0, atol=TOL)
        y_new = plants[0].step(u0)
        ctrl.store_input_output_measurement(u_current=u0, y_current=y_new, du_current=ctrl.get_du_value_at_step(n_step=0))
        orc.store_input_output_measurement(u0, y_new)
        ctrl.update_and_solve_data_driven_mpc()
        orc.update_and_solve_data_driven_mpc()
    np.testing.assert_allclose(ctrl.get_optimal_control_input_at_step(0), orc.get_optimal_control_input_at_step(0), rtol=0, atol=TOL)
    np.testing.assert_allclose(ctrl._b.alpha_prev[0].cpu().numpy(), orc.alpha_prev, rtol=0, atol=TOL)
    assert abs(float(ctrl._b.tracking_cost[0]) - orc.tracking_cost) <= 1e-6 * max(1.0, orc.tracking_cost)
