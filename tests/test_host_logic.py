"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, struct
layouts agree, parameter block values, config surface, sharding helpers, fail-loud behaviour."""

import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(build_lib):
    from data_driven_quad_control_b200 import _lib

    header = open(os.path.join(ROOT, "include", "ddqc.h")).read()
    declared = set(re.findall(r"\b(ddqc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(build_lib)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libddqc.so"
    loaded = _lib.load()  # also checks the ctypes struct sizes against ddqc_struct_size()
    assert loaded.ddqc_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define DDQC_ABI_VERSION (\d+)", header).group(1))
    assert loaded.ddqc_struct_size(99) == -1


def test_argument_validation_without_gpu(build_lib):
    """Entry points reject bad arguments before touching the device (no compute call is made)."""
    from data_driven_quad_control_b200 import _lib

    lib = _lib.load()
    P, S, O = _lib.EnvParams(), _lib.EnvState(), _lib.EnvOut()
    assert lib.ddqc_env_step(None, None, None, None, None, None, 16, None) == -1  # DDQC_E_NULL
    assert lib.ddqc_env_step(ctypes.byref(P), ctypes.byref(S), ctypes.byref(O), None, None, None, 0, None) == -2
    P.action_type, P.decimation = 7, 4
    assert lib.ddqc_env_step(ctypes.byref(P), ctypes.byref(S), ctypes.byref(O), None, None, None, 16, None) == -4
    cfg = _lib.MpcConfig()
    cfg.n, cfg.m, cfg.p, cfg.L, cfg.N = 6, 3, 3, 35, 350
    assert lib.ddqc_ddmpc_workspace_bytes(ctypes.byref(cfg), 4096) > 0
    # Gram cache: 24 header doubles + one workspace-slot image (R1 panel 34x5 blocks + trailing tri(29) blocks of 64) per controller
    assert lib.ddqc_ddmpc_gram_cache_bytes(ctypes.byref(cfg), 1) == 8 * (24 + 64 * (34 * 5 + 29 * 30 // 2))
    assert lib.ddqc_ddmpc_gram_cache_bytes(ctypes.byref(cfg), 4096) == 4096 * lib.ddqc_ddmpc_gram_cache_bytes(ctypes.byref(cfg), 1)
    assert lib.ddqc_ddmpc_gram_cache_bytes(ctypes.byref(cfg), -1) == -1
    assert lib.ddqc_ddmpc_solve(ctypes.byref(cfg), *([None] * 12), 16, None, None) == -1  # null buffers, with or without a cache
    assert lib.ddqc_ddmpc_store(ctypes.byref(cfg), None, None, None, None, 16, None, None) == -1
    cfg.alpha_reg_type = 3
    assert lib.ddqc_ddmpc_workspace_bytes(ctypes.byref(cfg), 4096) == -1
    # alpha_reg_type PREVIOUS needs v_sr = H alpha_sr (DdqcMpcAux), which the plain entry point cannot carry
    cfg.alpha_reg_type = 1
    assert lib.ddqc_ddmpc_workspace_bytes(ctypes.byref(cfg), 4096) > 0
    assert lib.ddqc_ddmpc_solve(ctypes.byref(cfg), *([ctypes.c_void_p(256)] * 12), 16, None, None) == -1
    aux = _lib.MpcAux()
    aux.u_ic = 256  # u_ic without y_ic
    assert lib.ddqc_ddmpc_solve_ex(ctypes.byref(cfg), *([ctypes.c_void_p(256)] * 12), 16, None, ctypes.byref(aux), None) == -1
    # dual recovery: workspace = 256 B + one (rp + 1) x ld matrix per slot (r = 6*42+1 = 253 -> rp = 256, ld = 272), 296 slots at most... 320
    assert lib.ddqc_ddmpc_recover_workspace_bytes(ctypes.byref(cfg), 1) == 256 + 8 * 257 * 272
    assert lib.ddqc_ddmpc_recover(ctypes.byref(cfg), *([None] * 15), 16, None) == -1
    assert lib.ddqc_ddmpc_hankel_apply(ctypes.byref(cfg), None, None, None, None, 16, None) == -1
    assert lib.ddqc_ddmpc_refresh_data(ctypes.byref(cfg), None, None, None, None, None, 1.0, None, None, 16, None) == -1
    cfg.alpha_reg_type, cfg.N = 0, 4000  # H would have > 512 rows? no: rows fixed by l; C large is fine
    assert lib.ddqc_ddmpc_workspace_bytes(ctypes.byref(cfg), 1) > 0
    cfg.L = 100  # r = 6*107+1 > kMaxR
    assert lib.ddqc_ddmpc_workspace_bytes(ctypes.byref(cfg), 1) == -1


def test_env_params_block():
    from data_driven_quad_control_b200.controllers.ctbr.ctbr_controller import build_ctbr_matrices
    from data_driven_quad_control_b200.envs.hover_env import make_env_params
    from data_driven_quad_control_b200.envs.hover_env_config import (
        EnvActionBounds,
        EnvActionType,
        EnvDroneParams,
        get_cfgs,
    )

    env_cfg, obs_cfg, rew_cfg, cmd_cfg = get_cfgs()
    mats = build_ctbr_matrices(EnvDroneParams.get())
    scales = {k: v * 0.04 for k, v in rew_cfg["reward_scales"].items()}
    P = make_env_params(env_cfg, obs_cfg, scales, rew_cfg["yaw_lambda"], cmd_cfg, EnvActionType.CTBR_FIXED_YAW, True,
                        False, 7, mixer=mats["mixer"], inv_sqrt_kf=float(mats["inv_sqrt_KF"]),
                        rate_pid_gains=[[10.0, 0.0, 0.1]] * 3)  # fmt: skip
    assert P.max_episode_length == 375 and P.min_hover_steps == 13 and P.decimation == 4
    assert P.flags == 3  # latency | auto target, no Euler needed at 180 deg thresholds
    assert abs(P.act_span[0] - 0.52974) < 1e-7 and P.act_lo[1] == -15.0 and P.act_span[2] == 30.0
    assert abs(P.inv_sqrt_kf - 56254.395) < 0.01 and abs(P.inv_step_dt - 25.0) < 1e-6
    np.testing.assert_allclose(list(P.mixer)[:4], [-1.2467909e-4, -1.2467905e-4, -2.1590723e-4, 0.24999937], rtol=2e-6)
    assert abs(EnvDroneParams.WEIGHT - 0.26487) < 1e-9 and abs(EnvActionBounds.MAX_THRUST - 0.52974) < 1e-9
    assert abs(EnvActionBounds.MIN_RPM - 2893.6858) < 1e-3 and abs(EnvActionBounds.MAX_RPM - 26043.1725) < 1e-3
    # the restated rigid-body constants equal those of the oracle bit for bit
    import drone_physics as dp

    Q = dp.PhysParams.cf2x(0.01)
    f32 = np.float32
    assert f32(P.dt_g) == Q.dt_g and f32(P.two_dt_over_m) == Q.two_dt_over_m and f32(P.half_dt) == Q.half_dt
    assert [f32(x) for x in P.k_w] == list(Q.k_w) and [f32(x) for x in P.b_w] == list(Q.b_w)
    assert [f32(x) for x in P.arm_x] == list(Q.arm_x) and [f32(x) for x in P.km_spin] == list(Q.km_spin)
    assert [f32(x) for x in P.cos_coef] == list(Q.cos_coef) and [f32(x) for x in P.sinc_coef] == list(Q.sinc_coef)
    # Euler angles are needed as soon as a threshold can fire or the yaw reward counts
    env_cfg["termination_if_roll_greater_than"] = 60
    assert make_env_params(env_cfg, obs_cfg, scales, -10.0, cmd_cfg, EnvActionType.ROTOR_RPMS, False, True, 0).flags & 4


def test_config_surface():
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, EnvState, get_cfgs
    from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type

    env_cfg, obs_cfg, rew_cfg, cmd_cfg = get_cfgs()
    assert list(rew_cfg["reward_scales"]) == ["target", "closeness", "hover_time", "smooth", "yaw", "angular", "crash"]
    assert EnvActionType.get_num_actions(EnvActionType.CTBR_FIXED_YAW) == 3
    assert EnvActionType.get_num_obs(EnvActionType.CTBR) == 17
    assert get_reward_cfg_by_action_type(EnvActionType.CTBR_FIXED_YAW)["reward_scales"]["yaw"] == 0.0
    assert get_reward_cfg_by_action_type(EnvActionType.ROTOR_RPMS)["reward_scales"]["smooth"] == -1e-4
    s = EnvState(*(torch.zeros(1, k) for k in (3, 4, 3, 3, 3)), torch.zeros(1, dtype=torch.int32), torch.zeros(1, 3))
    assert s.to_cpu().base_pos.device.type == "cpu" and s.ctbr_controller_state is None


@pytest.mark.reference
def test_config_matches_reference():
    import sys

    sys.path.insert(0, "/root/reference/src")
    from data_driven_quad_control.envs import hover_env_config as ref
    from data_driven_quad_control.learning.config.reward_config import get_reward_cfg_by_action_type as ref_rew

    from data_driven_quad_control_b200.envs import hover_env_config as ours
    from data_driven_quad_control_b200.learning.config.reward_config import get_reward_cfg_by_action_type

    assert ours.get_cfgs() == ref.get_cfgs()
    for at in ours.EnvActionType:
        assert get_reward_cfg_by_action_type(at) == ref_rew(ref.EnvActionType(at.value))
        assert ours.EnvActionType.get_num_obs(at) == ref.EnvActionType.get_num_obs(ref.EnvActionType(at.value))
    assert ours.EnvDroneParams.get() == ref.EnvDroneParams.get()
    assert ours.EnvCTBRControllerConfig.get() == ref.EnvCTBRControllerConfig.get()
    assert ours.EnvActionBounds.BASE_RPM == ref.EnvActionBounds.BASE_RPM


def test_math_utils_known_answers():
    """The known-answer vectors of the reference's tests/utility_tests/test_math_utils.py."""
    from data_driven_quad_control_b200.utilities.math_utils import (
        linear_interpolate,
        quaternion_to_matrix,
        yaw_from_quaternion,
        yaw_to_quaternion,
    )

    x = torch.tensor([0.0, 0.5, 1.0])
    torch.testing.assert_close(linear_interpolate(x, 0.0, 1.0, 10.0, 20.0), torch.tensor([10.0, 15.0, 20.0]))
    r = linear_interpolate(torch.tensor([0.25, 2.5]), torch.tensor([0.0, 0.0]), torch.tensor([1.0, 5.0]),
                           torch.tensor([100.0, 200.0]), torch.tensor([200.0, 500.0]))  # fmt: skip
    torch.testing.assert_close(r, torch.tensor([125.0, 350.0]))
    q = yaw_to_quaternion(torch.deg2rad(torch.tensor([0.0, 90.0])))
    torch.testing.assert_close(q, torch.tensor([[1.0, 0, 0, 0], [0.7071, 0, 0, 0.7071]]), rtol=0, atol=1e-4)
    R = quaternion_to_matrix(q)
    expected = torch.tensor([[[1.0, 0, 0], [0, 1, 0], [0, 0, 1]], [[0.0, -1, 0], [1, 0, 0], [0, 0, 1]]])
    torch.testing.assert_close(R, expected, rtol=0, atol=1e-6)
    quats = torch.tensor([[0.7071, 0, 0, 0.7071], [0.9239, 0.3827, 0, 0], [0, 0, 0.3827, 0.9239], [0.0990, -0.2391, 0.3696, 0.8924]])
    yaw = yaw_from_quaternion(quats)
    exp = torch.deg2rad(torch.tensor([90.0, 0.0, 180.0, 180.0]))
    torch.testing.assert_close(torch.sin(yaw), torch.sin(exp), rtol=0, atol=1e-4)
    torch.testing.assert_close(torch.cos(yaw), torch.cos(exp), rtol=0, atol=1e-4)


def test_no_cpu_fallback(build_lib):
    """Without a CUDA device every product class refuses to construct (no silent CPU path)."""
    from data_driven_quad_control_b200 import _lib
    from data_driven_quad_control_b200.envs.hover_env import HoverEnv
    from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, get_cfgs
    from data_driven_quad_control_b200.utilities.vectorized_pid_controller import VectorizedPIDController

    with pytest.raises(_lib.DdqcError):
        HoverEnv(4, *get_cfgs(), device="cpu", action_type=EnvActionType.CTBR)
    with pytest.raises(_lib.DdqcError):
        VectorizedPIDController(torch.ones(3, 3), 4, 0.04, device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(_lib.DdqcError):
            HoverEnv(4, *get_cfgs(), device="cuda", action_type=EnvActionType.CTBR)
    with pytest.raises(NotImplementedError):
        HoverEnv(4, *get_cfgs(), show_viewer=True)
    with pytest.raises(ValueError):
        VectorizedPIDController(torch.ones(3, 2), 4, 0.04)


def test_sharding_helpers():
    from data_driven_quad_control_b200.parallel import shard_range, shard_round_robin

    for n, w in ((1 << 20, 8), (4096, 3), (5, 8), (900, 8)):
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
        rr = sorted(i for r in range(w) for i in shard_round_robin(n, r, w))
        assert rr == list(range(n))
    with pytest.raises(ValueError):
        shard_range(10, 3, 2)


def test_closed_loop_oracle_rmse_and_early_stop():
    """oracle/ddmpc_eval_oracle.py follows controller_evaluation.py:345-450: RMSE = sqrt(mean_t sum (y - y_r)^2),
    early stop when |y - y_r| - d0 > max increment, n inputs applied per solve with n_n_mpc_step."""
    import ddmpc_eval_oracle as eo

    class Mock:  # same surface as the reference's tests/mocks.py:133-158
        n, m = 2, 1
        optimal_u = None

        def __init__(self):
            self.y_r = np.array([[1.0], [0.0]])
            self.solves, self.stored = 0, []

        def update_and_solve_data_driven_mpc(self):
            self.solves += 1
            self.optimal_u = np.array([[0.5], [0.25]])

        def get_optimal_control_input_at_step(self, n_step=0):
            return self.optimal_u[n_step]

        def get_du_value_at_step(self, n_step=0):
            return None

        def store_input_output_measurement(self, u_current, y_current, du_current):
            self.stored.append((float(u_current[0]), tuple(y_current)))

    ys = iter([np.array([0.0, 0.0]), np.array([0.5, 0.0]), np.array([1.0, 1.0]), np.array([1.0, 0.0])])
    c = Mock()
    rmse, u_sys, y_sys = eo.sim_control_loop(c, lambda u: next(ys), 4, 1.0, None, n_n_mpc_step=True)
    assert c.solves == 2 and [s[0] for s in c.stored] == [0.5, 0.25, 0.5, 0.25]
    assert rmse == pytest.approx(np.sqrt((1.0 + 0.25 + 1.0 + 0.0) / 4))
    ys = iter([np.array([0.0, 0.0]), np.array([-1.0, 0.0])])
    with pytest.raises(eo.MovedAway):
        eo.sim_control_loop(Mock(), lambda u: next(ys), 4, 1.0, 0.5)


def test_persistently_exciting_inputs_follow_the_reference_numpy_stream():
    """Same numpy calls in the same order as drone_initial_data_collection.py:117-127 of the reference, so the
    excitation and the generator state afterwards are identical to a single-env reference collection."""
    from data_driven_quad_control_b200.data_driven_mpc.utilities.drone_initial_data_collection import (
        persistently_exciting_inputs,
    )

    u_range = np.array([[-0.1, 0.1], [-0.5, 0.5], [-0.5, 0.5]])
    N, m, p, eps_max = 50, 3, 3, 0.002
    pe, w = persistently_exciting_inputs([np.random.default_rng(s) for s in (0, 7)], u_range, N, m, p, eps_max)
    for b, seed in enumerate((0, 7)):
        rng = np.random.default_rng(seed)
        u_ref = np.hstack([rng.uniform(u_range[i, 0], u_range[i, 1], (N, 1)) for i in range(m)])  # reference :117-122
        w_ref = eps_max * rng.uniform(-1.0, 1.0, (N, p))  # reference :125-127
        assert np.array_equal(pe[b], u_ref) and np.array_equal(w[b], w_ref)


def test_grid_search_bookkeeping_follows_the_reference_worker():
    """aggregate_combination == the sequential bookkeeping of parallel_worker.py:117-210 + controller_evaluation.py:
    120-228 (entries in order, setpoints in order, stop at the first failure, nanmean of per-entry means)."""
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.batched_grid_search import aggregate_combination
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.param_grid_search_config import (
        CtrlEvalStatus, DDMPCCombinationParams, DDMPCParameterGrid, combinations_from_grid,
    )  # fmt: skip

    comb = DDMPCCombinationParams(350, 6, 35, 50.0, 1e9, 1.0, 1e7)
    rmse = np.array([[0.2, 0.4, 0.6], [0.1, 0.3, 0.5]])
    status, res = aggregate_combination(comb, rmse, np.zeros((2, 3), bool))
    assert status == CtrlEvalStatus.SUCCESS and list(res)[:7] == list(comb._fields)
    assert res["average_RMSE"] == pytest.approx(np.mean([0.4, 0.3]))
    failed = np.array([[False, False, False], [False, True, False]])  # second entry fails on its second setpoint
    status, res = aggregate_combination(comb, rmse, failed)
    assert status == CtrlEvalStatus.FAILURE and list(res)[0] == "n_successful_runs" and res["n_successful_runs"] == 4
    assert res["average_RMSE"] == pytest.approx(np.mean([0.4, 0.1]))  # entry 2 averaged over its one completed run
    status, res = aggregate_combination(comb, rmse, np.array([[True, False, False], [False, False, False]]))
    assert res["n_successful_runs"] == 0 and np.isnan(res["average_RMSE"])  # nothing completed: NaN, later entries not run
    grid = DDMPCParameterGrid([350], [4, 6], [30, 35], [1e1, 5e1], [1e5], [1e-3], [1e3, 1e5])
    combos = combinations_from_grid(grid)
    assert len(combos) == 16 and combos[0] == (350, 4, 30, 1e1, 1e5, 1e-3, 1e3) and combos[1].lamb_sigma_s == 1e5 and combos[2].lamb_alpha == 5e1


def test_cache_states_keep_their_rate_pid_rows():
    """the per-env rows of the rate-PID state saved right after a data collection travel with the cached
    drone state (`row_of_state`) and are re-assembled for the evaluation batch (`stack_states`), instead of every
    evaluation run starting from a zeroed PID (a derivative kick the reference's runs do not have)."""
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.param_grid_search_config import (
        row_of_state,
        stack_states,
    )
    from data_driven_quad_control_b200.envs.hover_env_config import EnvState
    from data_driven_quad_control_b200.utilities.vectorized_pid_controller import VectorizedControllerState

    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    full = EnvState(base_pos=r(3, 3), base_quat=r(3, 4), base_lin_vel=r(3, 3), base_ang_vel=r(3, 3), commands=r(3, 3),
                    episode_length=torch.arange(3, dtype=torch.int32), last_actions=r(3, 3),
                    ctbr_controller_state=VectorizedControllerState(integral=r(3, 3), prev_error=r(3, 3)))  # fmt: skip
    rows = [row_of_state(full, b) for b in range(3)]
    assert rows[1].ctbr_controller_state.prev_error.shape == (1, 3)
    assert torch.equal(rows[1].ctbr_controller_state.prev_error[0], full.ctbr_controller_state.prev_error[1])
    # an evaluation batch: entry 2 three times, then entry 0
    batch = stack_states([rows[2], rows[2], rows[2], rows[0]])
    assert torch.equal(batch.ctbr_controller_state.integral, full.ctbr_controller_state.integral[[2, 2, 2, 0]])
    assert torch.equal(batch.base_pos, full.base_pos[[2, 2, 2, 0]])
    # states without a controller state (RPM action type) stack to None; a mixed list pads with zeros
    bare = row_of_state(EnvState(**{**full.__dict__, "ctbr_controller_state": None}), 0)
    assert stack_states([bare, bare]).ctbr_controller_state is None
    mixed = stack_states([rows[1], bare])
    assert torch.equal(mixed.ctbr_controller_state.prev_error[1], torch.zeros(3))
    # an explicit controller state still overrides
    z = VectorizedControllerState(integral=torch.zeros(2, 3), prev_error=torch.zeros(2, 3))
    assert stack_states([rows[1], rows[2]], z).ctbr_controller_state is z
