"""Data-driven MPC parameter grid search with every evaluation run in flight at once.

The reference (`parallel_grid_search.py:91-260`, `parallel_worker.py:60-240`, `controller_evaluation.py:40-230`)
hands each parameter combination to one of a handful of CPU worker processes; a worker walks the collected data
entries and, for each, the evaluation setpoints one after the other: a FRESH controller is built from the entry's
(u_N, y_N), the drone is restored to the state saved right after that collection, and the closed loop runs for
`eval_time_steps` steps.  Every such run is therefore independent of every other, and the only sequential
element is the bookkeeping "stop at the first failed run".  Here all runs of all combinations of one shape bucket
(N, n, L) - the QP shape fixes the kernel's shared-memory layout - are evaluated together: run r is controller r
on env r of one `HoverEnv`, its regularisation weights come from the per-controller `lamb` array, and the closed
loop is `sim_nonlinear_dd_mpc_control_loop_batched`.  The worker's stop-at-first-failure rule is applied afterwards
(`aggregate_combination`), so the reported status / n_successful_runs / average_RMSE are the ones the
reference's bookkeeping yields for the same per-run outcomes.

Multi-GPU: combinations are dealt round-robin to the ranks (`parallel.shard_round_robin`); each rank fills its
rows of an (n_combinations, 3) table [status, n_successful_runs, average_RMSE] and ONE all-reduce (sum) merges
them (`parallel.allreduce_cost_vector`).
"""

from __future__ import annotations

import copy
from typing import Any, Callable

import numpy as np
import torch

from data_driven_quad_control_b200 import parallel
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import (
    BatchedNonlinearDataDrivenMPC,
)
from data_driven_quad_control_b200.envs.hover_env import HoverEnv
from data_driven_quad_control_b200.utilities.vectorized_pid_controller import VectorizedControllerState

from .controller_evaluation import sim_nonlinear_dd_mpc_control_loop_batched
from .param_grid_search_config import (
    CtrlEvalStatus,
    DataDrivenCache,
    DDMPCCombinationParams,
    DDMPCEvaluationParams,
    DDMPCFixedParams,
    DDMPCParameterGrid,
    combinations_from_grid,
    stack_states,
)


def aggregate_combination(combination: DDMPCCombinationParams, rmse: np.ndarray, failed: np.ndarray,
                          reasons: list[list[str]] | None = None) -> tuple[CtrlEvalStatus, dict[str, Any]]:  # fmt: skip
    """Result of one combination from its per-run outcomes `rmse[entry, setpoint]`, `failed[entry, setpoint]`,
    following the worker's sequential bookkeeping (parallel_worker.py:117-210, controller_evaluation.py:120-228):
    entries in order, setpoints in order, stop at the first failed run; per entry the mean RMSE of its completed
    runs (NaN if none); overall the nanmean of the per-entry means."""
    n_entries, n_setpoints = rmse.shape
    per_entry: list[float] = []
    total_ok = 0
    for e in range(n_entries):
        ok_runs = 0
        while ok_runs < n_setpoints and not failed[e, ok_runs]:
            ok_runs += 1
        per_entry.append(float(np.mean(rmse[e, :ok_runs])) if ok_runs else float("nan"))
        total_ok += ok_runs
        if ok_runs < n_setpoints:
            avg = float(np.nanmean(per_entry)) if np.isfinite(per_entry).any() else float("nan")
            reason = reasons[e][ok_runs] if reasons else "Drone moved away from its target."
            return CtrlEvalStatus.FAILURE, {"n_successful_runs": total_ok, **combination._asdict(), "average_RMSE": avg,
                                            "failure_reason": reason}  # fmt: skip
    return CtrlEvalStatus.SUCCESS, {**combination._asdict(), "average_RMSE": float(np.nanmean(per_entry))}


def _weights(fixed: DDMPCFixedParams, n: int, L: int):
    ell = L + n + 1  # controller_creation.py:73-88
    return (np.kron(np.eye(ell), np.diag(fixed.Q_weights)), np.kron(np.eye(ell), np.diag(fixed.R_weights)),
            np.diag(fixed.S_weights))  # fmt: skip


def evaluate_bucket(make_env: Callable[[int], HoverEnv], cache: DataDrivenCache, combos: list[DDMPCCombinationParams],
                    fixed: DDMPCFixedParams, eval_params: DDMPCEvaluationParams, device) -> tuple[np.ndarray, np.ndarray]:  # fmt: skip
    """All runs of combinations that share (N, n, L).  Returns rmse, failed with shape (len(combos), entries,
    setpoints)."""
    N, n, L = combos[0].N, combos[0].n, combos[0].L
    entries = cache.N_to_entry_indices[N]
    setpoints = [np.asarray(s, dtype=np.float64).reshape(-1) for s in eval_params.eval_setpoints]
    nC, nE, nS = len(combos), len(entries), len(setpoints)
    B = nC * nE * nS
    run = lambda c, e, s: (c * nE + e) * nS + s  # noqa: E731
    u = np.empty((B, N, fixed.m))
    y = np.empty((B, N, fixed.p))
    y_r = np.empty((B, fixed.p))
    lam = np.empty((4, B))
    states = []
    for c, comb in enumerate(combos):
        for e, entry in enumerate(entries):
            for s, sp in enumerate(setpoints):
                r = run(c, e, s)
                u[r], y[r], y_r[r] = cache.u_N[entry], cache.y_N[entry], sp
                lam[:, r] = (comb.lamb_alpha, comb.lamb_sigma, comb.lamb_alpha_s, comb.lamb_sigma_s)
                states.append(cache.drone_state[entry])
    env = make_env(B)
    env.reset()
    idx = torch.arange(B, device=env.device)
    states = [s.to(env.device) for s in states]
    pid = None  # per-run rows of the rate-PID state saved right after the data collection (zero for caches without it)
    if not any(s.ctbr_controller_state is not None for s in states):
        pid = VectorizedControllerState(integral=torch.zeros((B, 3), device=env.device), prev_error=torch.zeros((B, 3), device=env.device))
    state = stack_states(states, pid)
    env.restore_from_state(idx, state)
    env.update_target_pos(idx, torch.from_numpy(y_r).to(env.device, torch.float32))
    Q, R, S = _weights(fixed, n, L)
    ctrl = BatchedNonlinearDataDrivenMPC(
        n=n, m=fixed.m, p=fixed.p, u=u, y=y, L=L, Q=Q, R=R, S=S, y_r=y_r, lamb_alpha=lam[0], lamb_sigma=lam[1], U=fixed.U,
        Us=fixed.Us, alpha_reg_type=fixed.alpha_reg_type, lamb_alpha_s=lam[2], lamb_sigma_s=lam[3],
        ext_out_incr_in=fixed.ext_out_incr_in, n_mpc_step=n if fixed.n_n_mpc_step else 1, device=device, solve_on_init=False,
    )  # fmt: skip
    d0 = np.linalg.norm(state.base_pos.cpu().numpy().astype(np.float64) - y_r, axis=1)  # controller_evaluation.py:138-150
    res = sim_nonlinear_dd_mpc_control_loop_batched(env, ctrl, eval_params.eval_time_steps, torch.from_numpy(d0),
                                                    eval_params.max_target_dist_increment, fixed.n_n_mpc_step)  # fmt: skip
    rmse = res.rmse.cpu().numpy().reshape(nC, nE, nS)
    failed = res.failed.cpu().numpy().reshape(nC, nE, nS)
    # the evaluation configuration disables the low-altitude termination; the reference's floor is not modelled here
    env.check_ground_plane(f"DD-MPC evaluation of {nC} combination(s)")
    env.close()
    return rmse, failed


def batched_dd_mpc_grid_search(make_env: Callable[[int], HoverEnv], cache: DataDrivenCache, fixed: DDMPCFixedParams,
                               eval_params: DDMPCEvaluationParams, param_grid: DDMPCParameterGrid, device="cuda",
                               max_runs_in_flight: int = 8192) -> dict[CtrlEvalStatus, list[dict[str, Any]]]:  # fmt: skip
    """Grid search over `param_grid`; returns the reference's `results` dictionary (status -> result dicts),
    identical on every rank."""
    combos = combinations_from_grid(param_grid)
    rank, world = parallel.rank_world()
    mine = parallel.shard_round_robin(len(combos), rank, world)
    table = torch.zeros((len(combos), 3), dtype=torch.float64, device=device)
    buckets: dict[tuple[int, int, int], list[int]] = {}
    for i in mine:
        buckets.setdefault((combos[i].N, combos[i].n, combos[i].L), []).append(i)
    nE_nS = eval_params.num_collections_per_N * len(eval_params.eval_setpoints)
    per_launch = max(1, max_runs_in_flight // nE_nS)
    for ids in buckets.values():
        for k in range(0, len(ids), per_launch):
            chunk = ids[k : k + per_launch]
            rmse, failed = evaluate_bucket(make_env, cache, [combos[i] for i in chunk], fixed, eval_params, device)
            for j, i in enumerate(chunk):
                status, result = aggregate_combination(combos[i], rmse[j], failed[j])
                avg = result["average_RMSE"]
                table[i] = torch.tensor([status.value, result.get("n_successful_runs", nE_nS), -1.0 if np.isnan(avg) else avg],
                                        dtype=torch.float64)  # fmt: skip
    flat = table.reshape(-1)
    table = parallel.allreduce_cost_vector(flat, list(range(flat.numel())), flat.numel()).reshape(len(combos), 3).cpu().numpy()
    results: dict[CtrlEvalStatus, list[dict[str, Any]]] = {CtrlEvalStatus.SUCCESS: [], CtrlEvalStatus.FAILURE: []}
    for comb, (status, n_ok, avg) in zip(combos, table):
        avg = float("nan") if avg < 0 else float(avg)
        if int(round(status)) == CtrlEvalStatus.SUCCESS.value:
            results[CtrlEvalStatus.SUCCESS].append({**comb._asdict(), "average_RMSE": avg})
        else:
            results[CtrlEvalStatus.FAILURE].append({"n_successful_runs": int(round(n_ok)), **comb._asdict(), "average_RMSE": avg,
                                                    "failure_reason": "Drone moved away from its target or the solve failed."})  # fmt: skip
    return results


def collect_data_driven_cache(env: HoverEnv, stabilizing_controller, target_pos: torch.Tensor, target_yaw: torch.Tensor,
                              input_bounds, u_range, N_values: list[int], num_collections_per_N: int,
                              np_random: np.random.Generator) -> DataDrivenCache:  # fmt: skip
    """`DataDrivenCache` with `num_collections_per_N` entries per trajectory length, all entries of one length
    collected at once (env b = entry b).  `np_random` seeds one child generator per entry
    (`np_random.spawn`), so entries are reproducible and independent."""
    from data_driven_quad_control_b200.data_driven_mpc.utilities.drone_initial_data_collection import (
        collect_initial_input_output_data_batched,
    )

    from .param_grid_search_config import row_of_state

    if env.num_envs != num_collections_per_N:
        raise ValueError("the collection env needs one drone per data entry (num_envs == num_collections_per_N)")
    cache = DataDrivenCache({}, {}, {}, {}, {})
    idx_all = torch.arange(env.num_envs, device=env.device)
    start = env.get_current_state(idx_all)
    entry = 0
    for N in N_values:
        env.restore_from_state(idx_all, copy.deepcopy(start))  # every length starts from the same stabilised hover
        stabilizing_controller.reset()  # initial_data_collection_cache.py:141: no position-PID state leaks between collections
        u_N, y_N = collect_initial_input_output_data_batched(
            env, stabilizing_controller, target_pos, target_yaw, torch.as_tensor(np.asarray(input_bounds)),
            torch.as_tensor(np.asarray(u_range)), N, np_randoms=np_random.spawn(env.num_envs),
        )  # fmt: skip
        after = env.get_current_state(idx_all)
        cache.N_to_entry_indices[N] = []
        for b in range(env.num_envs):
            cache.N_to_entry_indices[N].append(entry)
            cache.N[entry] = N
            cache.u_N[entry], cache.y_N[entry] = u_N[b].cpu().numpy(), y_N[b].cpu().numpy()
            cache.drone_state[entry] = row_of_state(after, b)
            entry += 1
    return cache
