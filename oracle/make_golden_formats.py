"""TEST INFRASTRUCTURE.  Golden files for SURVEY 8 f4 (report / trajectory formats), produced by the REFERENCE's own
code in the build container (needs /root/reference; third-party imports it cannot satisfy offline are stubbed:
`direct_data_driven_mpc...AlphaRegType` is an Enum with the members the reference uses,
grid_search_param_loader.py:24-28; matplotlib and the plotting helper are never called).

    python oracle/make_golden_formats.py     ->  tests/golden/ref_grid_search_report.txt
                                                 tests/golden/ref_control_trajectory.pkl
"""

import enum
import os
import pickle
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def install_reference_stubs() -> None:
    """Makes `data_driven_quad_control.data_driven_mpc.utilities.param_grid_search.results_writer` and
    `data_driven_quad_control.utilities.control_data_plotting` importable from /root/reference/src."""

    class AlphaRegType(enum.Enum):
        APPROXIMATED = 0
        PREVIOUS = 1
        ZERO = 2

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    mod("direct_data_driven_mpc")
    mod("direct_data_driven_mpc.nonlinear_data_driven_mpc_controller", AlphaRegType=AlphaRegType)
    mod("direct_data_driven_mpc.utilities")
    mod("direct_data_driven_mpc.utilities.visualization")
    mod("direct_data_driven_mpc.utilities.visualization.comparison_plot", plot_input_output_comparison=None)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mod("matplotlib")
        mod("matplotlib.pyplot")
        mod("matplotlib.axes", Axes=object)
        mod("matplotlib.figure", Figure=object)
    if "/root/reference/src" not in sys.path:
        sys.path.insert(0, "/root/reference/src")


def report_inputs(ns):
    """The same report inputs for either side; `ns` provides the parameter classes (reference's or this repo's)."""
    init = ns.DDMPCInitialDataCollectionParams(init_hover_pos=np.array([0.0, 0.0, 1.5]),
                                               u_range=np.array([[-0.1, 0.1], [-0.5, 0.5], [-0.5, 0.5]]))  # fmt: skip
    fixed = ns.DDMPCFixedParams(m=3, p=3, Q_weights=[1.0, 1.0, 1.0], R_weights=[0.1, 0.01, 0.01], S_weights=[50.0, 50.0, 200.0],
                                U=np.array([[0.0, 0.52974], [-1.5707963, 1.5707963], [-1.5707963, 1.5707963]]),
                                Us=np.array([[0.1, 0.45], [-0.5, 0.5], [-0.5, 0.5]]), alpha_reg_type=ns.AlphaRegType.APPROXIMATED,
                                ext_out_incr_in=False, n_n_mpc_step=True)  # fmt: skip
    ev = ns.DDMPCEvaluationParams(eval_time_steps=200, eval_setpoints=[np.array([[1.0], [1.0], [2.5]]), np.array([0.0, -1.0, 1.0])],
                                  max_target_dist_increment=0.5, num_collections_per_N=3)  # fmt: skip
    grid = ns.DDMPCParameterGrid(N=[350], n=[4, 6], L=[30, 35], lamb_alpha=[10.0, 50.0], lamb_sigma=[1e9],
                                 lamb_alpha_s=[1.0], lamb_sigma_s=[1e7, 1e5])  # fmt: skip
    S, F = ns.CtrlEvalStatus.SUCCESS, ns.CtrlEvalStatus.FAILURE
    base = dict(N=350, n=6, L=35, lamb_alpha=50.0, lamb_sigma=1e9, lamb_alpha_s=1.0, lamb_sigma_s=1e7)
    results = {
        S: [dict(base, average_RMSE=0.5062311, n_successful_runs=6, status="SUCCESS"),
            dict(base, n=4, average_RMSE=0.4681, n_successful_runs=6, status="SUCCESS"),
            dict(base, L=30, average_RMSE=0.61, n_successful_runs=6, status="SUCCESS")],
        F: [dict(base, lamb_alpha=10.0, average_RMSE=float("nan"), n_successful_runs=0, status="FAILURE"),
            dict(base, lamb_alpha=10.0, n=4, average_RMSE=0.9, n_successful_runs=4, status="FAILURE"),
            dict(base, lamb_sigma_s=1e5, average_RMSE=0.7, n_successful_runs=4, status="FAILURE"),
            dict(base, lamb_sigma_s=1e5, n=4, average_RMSE=1.25, n_successful_runs=5, status="FAILURE")],
    }  # fmt: skip
    return init, fixed, ev, grid, results


def trajectory_inputs():
    rng = np.random.default_rng(7)
    return dict(control_inputs=[rng.normal(size=(5, 3)), rng.normal(size=(5, 3)), rng.normal(size=(5, 3))],
                system_outputs=[rng.normal(size=(5, 3)) for _ in range(3)], system_setpoint=np.tile([1.0, 1.0, 2.5], (5, 1)))  # fmt: skip


def main() -> None:
    install_reference_stubs()
    import data_driven_quad_control.data_driven_mpc.utilities.param_grid_search.param_grid_search_config as cfg
    import data_driven_quad_control.data_driven_mpc.utilities.param_grid_search.results_writer as rw
    from data_driven_quad_control.utilities.control_data_plotting import ControlTrajectory
    from direct_data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType

    ns = types.SimpleNamespace(**{k: getattr(cfg, k) for k in dir(cfg) if k.startswith(("DDMPC", "CtrlEval"))}, AlphaRegType=AlphaRegType)
    init, fixed, ev, grid, results = report_inputs(ns)
    path = rw.write_results_to_file(GOLDEN, 3725.5, 10, init, fixed, ev, grid, results, file_name="ref_grid_search_report.txt")
    print("wrote", path)
    with open(os.path.join(GOLDEN, "ref_control_trajectory.pkl"), "wb") as f:
        pickle.dump(ControlTrajectory(**trajectory_inputs()), f, protocol=4)
    print("wrote", f.name)


if __name__ == "__main__":
    main()
