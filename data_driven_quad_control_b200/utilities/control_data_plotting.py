"""`ControlTrajectory`: the record the reference's comparison run produces and its plotting consumes
(utilities/control_data_plotting.py:15-19).  Plotting itself is out of scope here.

Pickle compatibility: the reference stores comparison results with `pickle.dump(trajectory)`
(comparison/controller_comparison.py) and `controller_comparison_plot.py:77` loads them back, so the pickle names the
class as `data_driven_quad_control.utilities.control_data_plotting.ControlTrajectory`.  `dump_control_trajectory`
writes exactly that global, so a file written here loads in a process that has only the reference installed;
`load_control_trajectory` reads files written by either side.
"""

from __future__ import annotations

import io
import pickle
import sys
import threading
import types
from dataclasses import dataclass

import numpy as np

REFERENCE_MODULE = "data_driven_quad_control.utilities.control_data_plotting"
_OWN_MODULE = __name__
_alias_lock = threading.Lock()


@dataclass
class ControlTrajectory:
    control_inputs: list[np.ndarray]  # per drone: (T, m) control inputs in physical units
    system_outputs: list[np.ndarray]  # per drone: (T, 3) true positions
    system_setpoint: np.ndarray  # (T, 3) target position at every step


class _ReferenceAlias:
    """While active, `ControlTrajectory` pickles under the reference's module path.  The alias modules are only put in
    `sys.modules` for names that are not already imported (a real reference install is never shadowed) and are removed
    again on exit."""

    def __enter__(self):
        _alias_lock.acquire()
        self._added: list[str] = []
        parts = REFERENCE_MODULE.split(".")
        for k in range(1, len(parts) + 1):
            name = ".".join(parts[:k])
            if name not in sys.modules:
                mod = types.ModuleType(name)
                mod.__path__ = []  # a package, so that the sub-module import machinery accepts it
                sys.modules[name] = mod
                self._added.append(name)
        leaf = sys.modules[REFERENCE_MODULE]
        self._had = getattr(leaf, "ControlTrajectory", None)
        leaf.ControlTrajectory = ControlTrajectory
        ControlTrajectory.__module__ = REFERENCE_MODULE
        return self

    def __exit__(self, *exc):
        ControlTrajectory.__module__ = _OWN_MODULE
        leaf = sys.modules.get(REFERENCE_MODULE)
        if leaf is not None and REFERENCE_MODULE not in self._added:
            if self._had is None:
                del leaf.ControlTrajectory
            else:
                leaf.ControlTrajectory = self._had
        for name in self._added:
            sys.modules.pop(name, None)
        _alias_lock.release()
        return False


def dumps_control_trajectory(trajectory: ControlTrajectory, protocol: int | None = None) -> bytes:
    with _ReferenceAlias():
        return pickle.dumps(trajectory, protocol=protocol)


def dump_control_trajectory(trajectory: ControlTrajectory, file) -> None:
    """`pickle.dump(trajectory, file)` in the reference's format (`file`: a path or a binary file object)."""
    data = dumps_control_trajectory(trajectory)
    if isinstance(file, str) or hasattr(file, "__fspath__"):
        with open(file, "wb") as f:
            f.write(data)
    else:
        file.write(data)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == "ControlTrajectory" and module in (REFERENCE_MODULE, _OWN_MODULE):
            return ControlTrajectory
        return super().find_class(module, name)


def load_control_trajectory(file) -> ControlTrajectory:
    """Loads a trajectory pickled by the reference (or by `dump_control_trajectory`) into this package's class; `file` is a
    path, a binary file object or the pickle bytes."""
    if isinstance(file, (bytes, bytearray)):  # the pickle itself
        return _Unpickler(io.BytesIO(file)).load()
    if isinstance(file, str) or hasattr(file, "__fspath__"):
        with open(file, "rb") as f:
            return _Unpickler(f).load()
    return _Unpickler(file).load()
