"""the grid-search report text and the `ControlTrajectory` pickle are the REFERENCE's formats.

* golden tests (run everywhere): the writer of this package reproduces, byte for byte (time stamp masked), the report the
  reference's own `results_writer.py` wrote for the same inputs (`tests/golden/ref_grid_search_report.txt`, made by
  `oracle/make_golden_formats.py`); a trajectory pickled by the reference loads here and vice versa.
* `reference`-marked tests (build container): the same comparisons against the reference's code run live.
"""

import os
import pickle
import re
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "oracle"))

_STAMP = re.compile(r"^Report generated at: .*$", re.M)


def _own_namespace():
    import data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.param_grid_search_config as cfg
    from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType

    return types.SimpleNamespace(**{k: getattr(cfg, k) for k in dir(cfg) if k.startswith(("DDMPC", "CtrlEval"))}, AlphaRegType=AlphaRegType)


def _own_report(tmp_path) -> str:
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.results_writer import write_results_to_file
    from make_golden_formats import report_inputs

    init, fixed, ev, grid, results = report_inputs(_own_namespace())
    path = write_results_to_file(str(tmp_path), 3725.5, 10, init, fixed, ev, grid, results, file_name="own.txt")
    return open(path).read()


def test_report_equals_the_reference_writers_output_byte_for_byte(tmp_path):
    want = open(os.path.join(GOLDEN, "ref_grid_search_report.txt")).read()
    got = _own_report(tmp_path)
    assert _STAMP.sub("STAMP", got) == _STAMP.sub("STAMP", want)
    assert _STAMP.search(got) and "average_RMSE=NaN" in got and "(3/7)" in got  # the masked line exists; NaN and counts are exercised


def test_default_report_file_name_and_empty_sections(tmp_path):
    from data_driven_quad_control_b200.data_driven_mpc.utilities.param_grid_search.results_writer import write_results_to_file
    from make_golden_formats import report_inputs

    init, fixed, ev, grid, _ = report_inputs(_own_namespace())
    path = write_results_to_file(str(tmp_path / "new_dir"), 0.25, 1, init, fixed, ev, grid, {})
    assert re.fullmatch(r"grid_search_results_\d{8}_\d{6}\.txt", os.path.basename(path))
    text = open(path).read()
    assert "Successful Results (0/0):\n  No results for any parameter combination.\n" in text
    assert "Failed Results (0/0):\n  No results for any parameter combination.\n" in text
    assert "Grid search duration: 0h 0m 0.25s" in text


def test_trajectory_pickled_by_the_reference_loads_here():
    from data_driven_quad_control_b200.utilities.control_data_plotting import ControlTrajectory, load_control_trajectory
    from make_golden_formats import trajectory_inputs

    traj = load_control_trajectory(os.path.join(GOLDEN, "ref_control_trajectory.pkl"))
    want = trajectory_inputs()
    assert isinstance(traj, ControlTrajectory)
    assert all(np.array_equal(a, b) for a, b in zip(traj.control_inputs, want["control_inputs"]))
    assert all(np.array_equal(a, b) for a, b in zip(traj.system_outputs, want["system_outputs"]))
    assert np.array_equal(traj.system_setpoint, want["system_setpoint"])


def test_trajectory_pickle_names_the_reference_class_and_leaves_no_alias_behind():
    import pickletools

    from data_driven_quad_control_b200.utilities.control_data_plotting import (REFERENCE_MODULE, ControlTrajectory,
                                                                                dumps_control_trajectory, load_control_trajectory)  # fmt: skip
    from make_golden_formats import trajectory_inputs

    had = REFERENCE_MODULE in sys.modules
    blob = dumps_control_trajectory(ControlTrajectory(**trajectory_inputs()), protocol=4)
    strings = [arg for op, arg, _ in pickletools.genops(blob) if op.name in ("SHORT_BINUNICODE", "BINUNICODE")]
    assert strings[:2] == [REFERENCE_MODULE, "ControlTrajectory"]
    assert (REFERENCE_MODULE in sys.modules) == had and ControlTrajectory.__module__.startswith("data_driven_quad_control_b200")
    # the reference's own pickle of the same record is the same byte string (same class path, same dataclass state)
    assert blob == open(os.path.join(GOLDEN, "ref_control_trajectory.pkl"), "rb").read()
    back = load_control_trajectory(blob)
    assert np.array_equal(back.system_setpoint, trajectory_inputs()["system_setpoint"])
    assert isinstance(pickle.loads(pickle.dumps(back)), ControlTrajectory)  # ordinary pickling still works


@pytest.mark.reference
def test_reference_writer_live_and_reference_side_unpickle(tmp_path):
    """The reference's results_writer.py, imported from /root/reference with the AlphaRegType stub, writes the same text
    as this package for the same inputs; the reference's control_data_plotting.ControlTrajectory unpickles a file
    written here (in a subprocess, so that the reference package is the only `data_driven_quad_control` around)."""
    from data_driven_quad_control_b200.utilities.control_data_plotting import ControlTrajectory, dump_control_trajectory
    from make_golden_formats import trajectory_inputs

    pkl = tmp_path / "traj.pkl"
    dump_control_trajectory(ControlTrajectory(**trajectory_inputs()), str(pkl))
    own = tmp_path / "own_report.txt"
    own.write_text(_own_report(tmp_path))
    code = f"""
import sys, pickle, re
sys.path.insert(0, {os.path.join(ROOT, 'oracle')!r})
import numpy as np, types
import make_golden_formats as m
m.install_reference_stubs()
import data_driven_quad_control.data_driven_mpc.utilities.param_grid_search.param_grid_search_config as cfg
import data_driven_quad_control.data_driven_mpc.utilities.param_grid_search.results_writer as rw
import data_driven_quad_control.utilities.control_data_plotting as cdp
from direct_data_driven_mpc.nonlinear_data_driven_mpc_controller import AlphaRegType
assert cdp.__file__.startswith('/root/reference/'), cdp.__file__
t = pickle.load(open({str(pkl)!r}, 'rb'))
assert type(t) is cdp.ControlTrajectory
w = m.trajectory_inputs()
assert np.array_equal(t.system_setpoint, w['system_setpoint']) and all(np.array_equal(a, b) for a, b in zip(t.control_inputs, w['control_inputs']))
ns = types.SimpleNamespace(**{{k: getattr(cfg, k) for k in dir(cfg) if k.startswith(('DDMPC', 'CtrlEval'))}}, AlphaRegType=AlphaRegType)
path = rw.write_results_to_file({str(tmp_path)!r}, 3725.5, 10, *m.report_inputs(ns), file_name='ref_live.txt')
mask = lambda s: re.sub(r'^Report generated at: .*$', 'STAMP', s, flags=re.M)
assert mask(open(path).read()) == mask(open({str(own)!r}).read())
print('OK')
"""
    out = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]
