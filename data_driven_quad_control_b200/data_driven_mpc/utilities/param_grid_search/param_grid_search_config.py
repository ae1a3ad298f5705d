"""Parameter containers of the data-driven MPC grid search: same names and fields as the reference's
`data_driven_mpc/utilities/param_grid_search/param_grid_search_config.py:13-136`, so that its YAML loader,
report writer and callers map one to one.  (The multiprocessing bookkeeping types `EnvSimInfo` /
`EnvResetSignal` have no counterpart: there are no worker processes or queues here.)"""

from __future__ import annotations

from enum import Enum
from typing import NamedTuple

import numpy as np
import torch

from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType
from data_driven_quad_control_b200.envs.hover_env_config import EnvState


class DDMPCInitialDataCollectionParams(NamedTuple):
    init_hover_pos: np.ndarray
    u_range: np.ndarray


class DDMPCFixedParams(NamedTuple):
    m: int
    p: int
    Q_weights: list[float]
    R_weights: list[float]
    S_weights: list[float]
    U: np.ndarray
    Us: np.ndarray
    alpha_reg_type: AlphaRegType
    ext_out_incr_in: bool
    n_n_mpc_step: bool


class DDMPCEvaluationParams(NamedTuple):
    eval_time_steps: int
    eval_setpoints: list[np.ndarray]
    max_target_dist_increment: float
    num_collections_per_N: int


class DDMPCCombinationParams(NamedTuple):
    N: int
    n: int
    L: int
    lamb_alpha: float
    lamb_sigma: float
    lamb_alpha_s: float
    lamb_sigma_s: float


class DDMPCParameterGrid(NamedTuple):
    N: list[int]
    n: list[int]
    L: list[int]
    lamb_alpha: list[float]
    lamb_sigma: list[float]
    lamb_alpha_s: list[float]
    lamb_sigma_s: list[float]


class DataDrivenCache(NamedTuple):
    """Initial input-output measurements per data entry and the drone state right after each collection
    (reference :91-110).  `u_N` / `y_N` values are float64 arrays (N, m) / (N, p); `drone_state` one-row
    `EnvState`s."""

    N_to_entry_indices: dict[int, list[int]]
    N: dict[int, int]
    u_N: dict[int, np.ndarray]
    y_N: dict[int, np.ndarray]
    drone_state: dict[int, EnvState]


class CtrlEvalStatus(Enum):
    SUCCESS = 0
    FAILURE = 1


def combinations_from_grid(param_grid: DDMPCParameterGrid) -> list[DDMPCCombinationParams]:
    """All combinations in the reference's enumeration order (`itertools.product` over the grid fields in
    declaration order: N, n, L, lamb_alpha, lamb_sigma, lamb_alpha_s, lamb_sigma_s)."""
    import itertools

    return [DDMPCCombinationParams(*c) for c in itertools.product(*[getattr(param_grid, f) for f in param_grid._fields])]


def row_of_state(state: EnvState, row: int) -> EnvState:
    """One-row copy of a multi-env `EnvState`, INCLUDING that env's row of the rate-PID state: the reference saves the
    controller state right after a data collection and restores it before every evaluation run
    (hover_env.py:680-695, 737-741), so a run starts with the `prev_error` / `integral` the collection left behind and
    not with a derivative kick from zero."""
    from data_driven_quad_control_b200.utilities.vectorized_pid_controller import VectorizedControllerState

    pick = lambda t: t[row : row + 1].clone()  # noqa: E731
    pid = state.ctbr_controller_state
    if pid is not None:
        pid = VectorizedControllerState(integral=pick(pid.integral), prev_error=pick(pid.prev_error))
    return EnvState(base_pos=pick(state.base_pos), base_quat=pick(state.base_quat), base_lin_vel=pick(state.base_lin_vel),
                    base_ang_vel=pick(state.base_ang_vel), commands=pick(state.commands),
                    episode_length=pick(state.episode_length), last_actions=pick(state.last_actions),
                    ctbr_controller_state=pid)  # fmt: skip


def stack_states(states: list[EnvState], ctbr_state=None) -> EnvState:
    """Concatenates one-row states into the state of a batch.  The rate-PID rows are concatenated too (rows of states
    that carry none are zero); an explicit `ctbr_state` overrides them."""
    from data_driven_quad_control_b200.utilities.vectorized_pid_controller import VectorizedControllerState

    cat = lambda name: torch.cat([getattr(s, name) for s in states], dim=0)  # noqa: E731
    if ctbr_state is None and any(s.ctbr_controller_state is not None for s in states):
        ref = states[0].base_pos
        zero = lambda s: torch.zeros((s.base_pos.shape[0], 3), dtype=torch.float32, device=ref.device)  # noqa: E731
        part = lambda s, f: getattr(s.ctbr_controller_state, f).to(ref.device, torch.float32) if s.ctbr_controller_state is not None else zero(s)  # noqa: E731
        ctbr_state = VectorizedControllerState(integral=torch.cat([part(s, "integral") for s in states], dim=0),
                                               prev_error=torch.cat([part(s, "prev_error") for s in states], dim=0))  # fmt: skip
    return EnvState(base_pos=cat("base_pos"), base_quat=cat("base_quat"), base_lin_vel=cat("base_lin_vel"),
                    base_ang_vel=cat("base_ang_vel"), commands=cat("commands"), episode_length=cat("episode_length"),
                    last_actions=cat("last_actions"), ctbr_controller_state=ctbr_state)  # fmt: skip
