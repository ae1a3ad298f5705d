"""Collective-thrust / body-rate (CTBR) controller backed by `ddqc_ctbr_compute`.

Host mirror of the reference's `controllers/ctbr/ctbr_controller.py`
(`DroneCTBRController.__init__` :24-129, allocation matrix :131-179, mixer :181-219,
`compute` :221-264, `reset/get_state/load_state` :266-292).  The rate PID, the 4x4 mixer
product, the thrust clamp and the RPM conversion run in one CUDA kernel.

The allocation matrix, its pseudo-inverse and the mixer are built once on the HOST with the
same fp32 torch calls as the reference and handed to the kernel as 16 floats: float32
`torch.linalg.pinv` is device dependent in its last bits, and the mixer must be identical for
the kernel, the fused env step and the test oracle.
"""

from __future__ import annotations

import ctypes as C

import torch

from data_driven_quad_control_b200 import _lib
from data_driven_quad_control_b200.drone_config.drone_params import DroneParams
from data_driven_quad_control_b200.utilities.vectorized_pid_controller import (
    VectorizedControllerState,
    VectorizedPIDController,
)

from .ctbr_controller_config import CTBRControllerConfig


def build_ctbr_matrices(drone_params: DroneParams) -> dict[str, torch.Tensor]:
    """Allocation matrix A (4 x n_rotors), pinv(A), inertia J and mixer = pinv(A) @ diag(J, 1),
    all fp32 on the CPU.  Rows of A: roll torque -l sin(theta_i), pitch torque -l cos(theta_i),
    yaw torque (KM/KF) spin_i, total thrust 1."""
    phys, rotor = drone_params["drone_physical_params"], drone_params["drone_rotor_params"]
    f32 = dict(dtype=torch.float32, device="cpu")
    KF = torch.as_tensor(rotor["kf"], **f32)
    KM = torch.as_tensor(rotor["km"], **f32)
    i = phys["inertia"]
    J = torch.tensor([[i["Jxx"], i["Jxy"], i["Jxz"]], [i["Jxy"], i["Jyy"], i["Jyz"]], [i["Jxz"], i["Jyz"], i["Jzz"]]], **f32)
    theta = torch.deg2rad(torch.as_tensor(rotor["rotor_angles_deg"], **f32))
    spin = torch.as_tensor(rotor["rotor_spin_directions"], dtype=torch.int, device="cpu")
    arm = rotor["arm_length"]
    A = torch.vstack([-torch.sin(theta) * arm, -torch.cos(theta) * arm, KM / KF * spin, torch.ones_like(theta)])
    A_inv = torch.linalg.pinv(A)
    J_exp = torch.eye(4, **f32)
    J_exp[:3, :3] = J
    mixer = A_inv @ J_exp
    return {"KF": KF, "KM": KM, "J": J, "A": A, "A_inv": A_inv, "mixer": mixer, "inv_sqrt_KF": torch.rsqrt(KF),
            "rotor_angles": theta, "rotor_spin_directions": spin}


class DroneCTBRController:
    def __init__(
        self,
        drone_params: DroneParams,
        controller_config: CTBRControllerConfig,
        dt: float,
        num_envs: int,
        device: torch.device | str = "cuda",
    ):
        self._lib = _lib.load()
        self.num_envs = num_envs
        self.device = _lib.require_cuda(device)
        self.dt = dt

        mats = build_ctbr_matrices(drone_params)
        self.inertia = drone_params["drone_physical_params"]["inertia"]
        self.arm_length = drone_params["drone_rotor_params"]["arm_length"]
        for name in ("KF", "KM", "J", "A", "A_inv", "mixer", "inv_sqrt_KF", "rotor_angles", "rotor_spin_directions"):
            setattr(self, name, mats[name].to(self.device))
        self._mixer_host = (C.c_float * 16)(*mats["mixer"].flatten().tolist())
        self._inv_sqrt_kf_host = float(mats["inv_sqrt_KF"])

        gains = controller_config["ctbr_controller_params"]["rate_pid_gains"]
        self.rate_pid_params = torch.as_tensor(gains, dtype=torch.float, device=self.device)
        self._gains_host = (C.c_float * 9)(*[float(g) for row in gains for g in row])
        self.rate_pid_controller = VectorizedPIDController(
            pid_params=self.rate_pid_params, num_envs=num_envs, dt=dt, device=self.device
        )

    def compute(
        self, rate_measurements: torch.Tensor, rate_setpoints: torch.Tensor, thrust_setpoints: torch.Tensor
    ) -> torch.Tensor:
        """(N,3) measured rates, (N,3) rate setpoints, (N,) thrust setpoints -> (N,4) rotor RPMs."""
        dev, N = self.device, self.num_envs
        ms = rate_measurements.to(device=dev, dtype=torch.float32)
        sp = rate_setpoints.to(device=dev, dtype=torch.float32)
        th = thrust_setpoints.to(device=dev, dtype=torch.float32)
        if tuple(ms.shape) != (N, 3) or tuple(sp.shape) != (N, 3) or tuple(th.shape) != (N,):
            raise ValueError("expected rate_measurements (N,3), rate_setpoints (N,3), thrust_setpoints (N,)")
        rpm = torch.empty((N, 4), dtype=torch.float32, device=dev)
        integ, prev = self.rate_pid_controller._state_ptrs()
        with _lib.device_guard(dev):
            rc = self._lib.ddqc_ctbr_compute(
                ms.data_ptr(), ms.stride(0), ms.stride(1), sp.data_ptr(), sp.stride(0), sp.stride(1),
                th.data_ptr(), th.stride(0), integ, prev, self._mixer_host, self._gains_host,
                self.dt, self._inv_sqrt_kf_host, rpm.data_ptr(), N, _lib.stream_ptr(dev),
            )  # fmt: skip
        _lib.check(rc, "ddqc_ctbr_compute")
        return rpm

    def reset(self) -> None:
        self.rate_pid_controller.reset()

    def get_state(self) -> VectorizedControllerState:
        return self.rate_pid_controller.get_state()

    def load_state(self, state: VectorizedControllerState) -> None:
        self.rate_pid_controller.load_state(state)
