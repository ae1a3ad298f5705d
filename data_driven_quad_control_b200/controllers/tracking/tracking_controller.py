"""SE(3)-style PID tracking controller backed by `ddqc_tracking_compute`.

Host mirror of the reference's `controllers/tracking/tracking_controller.py`
(`DroneTrackingController.__init__` :30-88, `compute` :90-164, `reset` :166-173): position PID
-> gravity compensation -> thrust along the current attitude -> desired attitude from the
force direction and the yaw setpoint -> attitude error e_R -> body-rate command, fused in one
kernel that returns (N,4) CTBR commands [thrust, w_x, w_y, w_z].
"""

from __future__ import annotations

import ctypes as C

import torch

from data_driven_quad_control_b200 import _lib
from data_driven_quad_control_b200.utilities.vectorized_pid_controller import VectorizedPIDController

from .tracking_controller_config import TrackingControllerConfig, TrackingCtrlDroneState


class DroneTrackingController:
    def __init__(
        self,
        drone_mass: float,
        controller_config: TrackingControllerConfig,
        dt: float,
        num_envs: int,
        device: torch.device | str = "cuda",
    ):
        self._lib = _lib.load()
        self.device = _lib.require_cuda(device)
        self.num_envs = num_envs
        self.dt = dt
        params = controller_config["tracking_controller_params"]
        self.pos_pid_params = torch.as_tensor(params["pos_pid_gains"], dtype=torch.float, device=self.device)
        self._gains_host = (C.c_float * 9)(*[float(g) for row in params["pos_pid_gains"] for g in row])
        self.kR = params["kR"]
        self.drone_mass = drone_mass
        self.controller = VectorizedPIDController(
            pid_params=self.pos_pid_params, num_envs=num_envs, dt=dt, device=self.device
        )

    def compute(self, state_measurement: TrackingCtrlDroneState, state_setpoint: TrackingCtrlDroneState) -> torch.Tensor:
        dev, N = self.device, self.num_envs
        cast = lambda t: t.to(device=dev, dtype=torch.float32)  # noqa: E731
        X, Q = cast(state_measurement.X), cast(state_measurement.Q)
        Xs, Qs = cast(state_setpoint.X), cast(state_setpoint.Q)
        if tuple(X.shape) != (N, 3) or tuple(Xs.shape) != (N, 3) or tuple(Q.shape) != (N, 4) or tuple(Qs.shape) != (N, 4):
            raise ValueError("expected positions (N,3) and quaternions (N,4)")
        out = torch.empty((N, 4), dtype=torch.float32, device=dev)
        integ, prev = self.controller._state_ptrs()
        with _lib.device_guard(dev):
            rc = self._lib.ddqc_tracking_compute(
                X.data_ptr(), X.stride(0), X.stride(1), Q.data_ptr(), Q.stride(0), Q.stride(1),
                Xs.data_ptr(), Xs.stride(0), Xs.stride(1), Qs.data_ptr(), Qs.stride(0), Qs.stride(1),
                integ, prev, self._gains_host, self.dt, float(self.drone_mass), float(self.kR),
                out.data_ptr(), N, _lib.stream_ptr(dev),
            )  # fmt: skip
        _lib.check(rc, "ddqc_tracking_compute")
        return out

    def reset(self) -> None:
        self.controller.reset()
