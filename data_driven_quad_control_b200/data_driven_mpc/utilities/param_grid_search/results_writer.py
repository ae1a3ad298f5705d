"""Text report of a data-driven MPC parameter grid search, in the layout of the reference's
`data_driven_mpc/utilities/param_grid_search/results_writer.py:25-240` (asserted line by line by its
`tests/data_driven_mpc_tests/test_dd_mpc_param_grid_search.py:161-188`), so that whoever parses the
reference's reports parses these."""

from __future__ import annotations

import math
import os
from datetime import datetime
from typing import Any, Callable, NamedTuple, TextIO

from .param_grid_search_config import (
    CtrlEvalStatus,
    DDMPCEvaluationParams,
    DDMPCFixedParams,
    DDMPCInitialDataCollectionParams,
    DDMPCParameterGrid,
)

_ARRAY_KEYS = ("u_range", "U", "Us", "eval_setpoints")


def get_num_combinations_from_grid(param_grid: DDMPCParameterGrid) -> int:
    return math.prod(len(getattr(param_grid, name)) for name in param_grid._fields)


def format_result_dict(result: dict[str, Any]) -> str:
    def show(v):
        if isinstance(v, float) and math.isnan(v):
            return "NaN"
        return f'"{v}"' if isinstance(v, str) else v

    return ", ".join(f"{k}={show(v)}" for k, v in result.items())


def write_dd_mpc_grid_search_params_to_file(f: TextIO, dd_mpc_params: NamedTuple) -> None:
    for key, value in dd_mpc_params._asdict().items():
        if key == "alpha_reg_type":
            f.write(f"  {key}: {value.name}\n")
        elif key in _ARRAY_KEYS:
            f.write(f"  {key}:\n")
            for arr in value:
                f.write(f"    - {arr.flatten().tolist()}\n")
        else:
            f.write(f"  {key}: {value}\n")


def write_result_section(f: TextIO, title: str, result_list: list[dict[str, Any]], total_searches: int,
                         sort_key: Callable[[dict[str, Any]], Any], reverse: bool = False) -> None:  # fmt: skip
    ordered = sorted(result_list, key=sort_key, reverse=reverse)
    f.write(f"{title} ({len(ordered)}/{total_searches}):\n")
    if not ordered:
        f.write("  No results for any parameter combination.\n")
    for result in ordered:
        f.write(f"  {format_result_dict(result)}\n")


def write_results_to_file(output_dir: str, elapsed_time: float, num_processes: int,
                          init_data_collection_params: DDMPCInitialDataCollectionParams, fixed_params: DDMPCFixedParams,
                          eval_params: DDMPCEvaluationParams, param_grid: DDMPCParameterGrid,
                          results: dict[CtrlEvalStatus, list[dict[str, Any]]], file_name: str | None = None) -> str:  # fmt: skip
    """`num_processes` is the number of evaluation runs that were in flight at once (the reference's worker
    count); everything else as in the reference."""
    os.makedirs(output_dir, exist_ok=True)
    now = datetime.now()
    if file_name is None:
        file_name = f"grid_search_results_{now.strftime('%Y%m%d_%H%M%S')}.txt"
    path = os.path.join(output_dir, file_name)
    hours, rest = divmod(elapsed_time, 3600)
    minutes, seconds = divmod(rest, 60)
    n_comb = get_num_combinations_from_grid(param_grid)
    runs_per_comb = len(eval_params.eval_setpoints) * eval_params.num_collections_per_N
    total = sum(len(v) for v in results.values())
    with open(path, "w") as f:
        f.write(f"Report generated at: {now.strftime('%Y-%m-%d %H:%M:%S')}\n\n")
        f.write(f"Grid search duration: {int(hours)}h {int(minutes)}m {seconds:.2f}s\n")
        f.write(f"Number of parallel processes: {num_processes}\n\n")
        for title, params in (("Initial data collection parameters:", init_data_collection_params),
                              ("Fixed parameters:", fixed_params), ("Evaluation parameters:", eval_params),
                              ("Grid Search conducted over the following parameters:", param_grid)):  # fmt: skip
            f.write(title + "\n")
            write_dd_mpc_grid_search_params_to_file(f, params)
            f.write("\n")
        f.write("Grid Search summary:\n")
        f.write(f"  Total parameter combinations: {n_comb}\n")
        f.write(f"  Total evaluation runs: {n_comb * runs_per_comb}\n")
        f.write(f"  Evaluation runs per combination: {runs_per_comb}\n\n")
        write_result_section(f, "Successful Results", results.get(CtrlEvalStatus.SUCCESS, []), total,
                             sort_key=lambda r: r["average_RMSE"])  # fmt: skip
        f.write("\n")
        write_result_section(f, "Failed Results", results.get(CtrlEvalStatus.FAILURE, []), total,
                             sort_key=lambda r: (r["n_successful_runs"], -r["average_RMSE"]), reverse=True)  # fmt: skip
    return path
