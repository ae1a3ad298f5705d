"""Vectorized PID controller backed by the `ddqc_pid_compute` CUDA kernel.

Host mirror of the reference's `utilities/vectorized_pid_controller.py`
(`VectorizedControllerState` :8-22, `VectorizedPIDController` :25-158): same constructor,
attributes (`kp`, `ki`, `kd`, `integral`, `prev_error`, `dt`, `num_envs`, `num_ctrls`) and
methods (`compute`, `reset`, `get_state`, `load_state`).  Differences, by design: the state
tensors are updated in place by the kernel (the reference rebinds `prev_error` to a fresh
tensor every call) and the controller only exists on CUDA devices.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from data_driven_quad_control_b200 import _lib


@dataclass
class VectorizedControllerState:
    integral: torch.Tensor
    prev_error: torch.Tensor

    def clone(self) -> "VectorizedControllerState":
        return VectorizedControllerState(self.integral.clone(), self.prev_error.clone())

    def to(self, device: torch.device | str) -> "VectorizedControllerState":
        return VectorizedControllerState(self.integral.to(device), self.prev_error.to(device))


def _as_f32(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    if t.dtype != torch.float32 or t.device != device:
        t = t.to(device=device, dtype=torch.float32)
    return t


class VectorizedPIDController:
    def __init__(self, pid_params: torch.Tensor, num_envs: int, dt: float, device: torch.device | str = "cuda"):
        if pid_params.ndim != 2 or pid_params.shape[1] != 3:
            raise ValueError(
                "PID params must have shape (`num_ctrls`, 3), representing the P, I, and D gains "
                "for each of the `num_ctrls` controllers."
            )
        self._lib = _lib.load()
        self.device = _lib.require_cuda(device)
        self.num_envs = num_envs
        self.dt = dt
        self.num_ctrls = int(pid_params.shape[0])
        if not 1 <= self.num_ctrls <= 8:
            raise ValueError("VectorizedPIDController supports 1..8 controllers per env")
        gains = pid_params.detach().to("cpu", torch.float32).contiguous()
        self._gains_host = (C.c_float * (3 * self.num_ctrls))(*gains.flatten().tolist())
        params_dev = _as_f32(pid_params.detach(), self.device)
        self.kp = params_dev[:, 0].expand(num_envs, -1)
        self.ki = params_dev[:, 1].expand(num_envs, -1)
        self.kd = params_dev[:, 2].expand(num_envs, -1)
        self.integral = torch.zeros((num_envs, self.num_ctrls), dtype=torch.float32, device=self.device)
        self.prev_error = torch.zeros_like(self.integral)

    def _state_ptrs(self) -> tuple[int, int]:
        # load_state() may have bound arbitrary tensors: the kernel needs dense (N, k) fp32 storage
        if not (self.integral.is_contiguous() and self.integral.dtype == torch.float32 and self.integral.device == self.device):
            self.integral = _as_f32(self.integral, self.device).contiguous()
        if not (self.prev_error.is_contiguous() and self.prev_error.dtype == torch.float32 and self.prev_error.device == self.device):
            self.prev_error = _as_f32(self.prev_error, self.device).contiguous()
        return self.integral.data_ptr(), self.prev_error.data_ptr()

    def compute(self, setpoints: torch.Tensor, measurements: torch.Tensor) -> torch.Tensor:
        sp = _as_f32(setpoints, self.device)
        ms = _as_f32(measurements, self.device)
        shape = (self.num_envs, self.num_ctrls)
        if tuple(sp.shape) != shape or tuple(ms.shape) != shape:
            raise ValueError(f"setpoints / measurements must have shape {shape}")
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        integ, prev = self._state_ptrs()
        with _lib.device_guard(self.device):
            rc = self._lib.ddqc_pid_compute(
                sp.data_ptr(), sp.stride(0), sp.stride(1), ms.data_ptr(), ms.stride(0), ms.stride(1),
                integ, prev, self._gains_host, self.dt, out.data_ptr(), self.num_envs, self.num_ctrls,
                _lib.stream_ptr(self.device),
            )  # fmt: skip
        _lib.check(rc, "ddqc_pid_compute")
        return out

    def reset(self) -> None:
        self.integral.zero_()
        self.prev_error.zero_()

    def get_state(self) -> VectorizedControllerState:
        return VectorizedControllerState(self.integral.clone(), self.prev_error.clone())

    def load_state(self, state: VectorizedControllerState) -> None:
        self.integral = state.integral
        self.prev_error = state.prev_error
