"""ctypes binding of libddqc.so (C ABI declared in include/ddqc.h).

The structures below mirror the C structs field for field; `load()` checks their sizes
against `ddqc_struct_size()` so that a header / binding mismatch fails loudly at import time
of the product path.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DDQC_LIB", os.path.join(_HERE, "libddqc.so"))  # DDQC_LIB: kernel-variant experiments

NUM_REWARDS = 7
STAT_COPIES = 32
EP_RING = 1024

ABI_VERSION = 3  # include/ddqc.h: DDQC_ABI_VERSION

ACT_ROTOR_RPMS, ACT_CTBR, ACT_CTBR_FIXED_YAW = 0, 1, 2
F_ACTION_LATENCY, F_AUTO_TARGET, F_NEED_EULER, F_INDEPENDENT_ENVS, F_GROUND_WATCH = 1, 2, 4, 8, 16

f32, i32, u32, u64, f64 = C.c_float, C.c_int32, C.c_uint32, C.c_uint64, C.c_double
vp = C.c_void_p


class EnvParams(C.Structure):
    _fields_ = [
        ("action_type", i32),
        ("decimation", i32),
        ("max_episode_length", i32),
        ("min_hover_steps", i32),
        ("flags", u32),
        ("clip_actions", f32),
        ("act_lo", f32 * 4),
        ("act_span", f32 * 4),
        ("step_dt", f32),
        ("inv_step_dt", f32),
        ("pid_kp", f32 * 3),
        ("pid_ki", f32 * 3),
        ("pid_kd", f32 * 3),
        ("mixer", f32 * 16),
        ("inv_sqrt_kf", f32),
        ("dt", f32),
        ("half_dt", f32),
        ("dt_g", f32),
        ("kf", f32),
        ("dt_over_m", f32),
        ("two_dt_over_m", f32),
        ("k_w", f32 * 3),
        ("b_w", f32 * 3),
        ("arm_x", f32 * 4),
        ("arm_y", f32 * 4),
        ("km_spin", f32 * 4),
        ("cos_coef", f32 * 5),
        ("sinc_coef", f32 * 5),
        ("actuator_noise_std", f32),
        ("min_rpm", f32),
        ("max_rpm", f32),
        ("obs_noise_std", f32),
        ("term_pitch", f32),
        ("term_roll", f32),
        ("term_x", f32),
        ("term_y", f32),
        ("term_z", f32),
        ("term_ground", f32),
        ("term_ang_vel", f32),
        ("term_lin_vel", f32),
        ("init_pos", f32 * 3),
        ("init_quat", f32 * 4),
        ("inv_init_quat", f32 * 4),
        ("at_target_threshold", f32),
        ("at_target_threshold_sq", f32),
        ("inv_at_target_threshold_sq", f32),
        ("inv_pi_lit", f32),
        ("inv_180", f32),
        ("cmd_lo", f32 * 3),
        ("cmd_span", f32 * 3),
        ("obs_scale_pos", f32),
        ("obs_scale_lin", f32),
        ("obs_scale_ang", f32),
        ("rew_scale", f32 * NUM_REWARDS),
        ("yaw_lambda", f32),
        ("episode_length_s", f32),
        ("seed", u64),
    ]


class EnvCtl(C.Structure):
    _fields_ = [
        ("rng_event", u64),
        ("step_count", u64),
        ("total_resets", u64),
        ("total_timeouts", u64),
        ("total_env_steps", u64),
        ("total_reward", f64),
        ("total_episode_sums", f64 * NUM_REWARDS),
        ("total_below_ground", u64),
        ("reserved0", u64),
        ("blocks_done", u32),
        ("pid_reset_pending", u32),
        ("n_reset_last", u32),
        ("fixup_count", u32),
        ("ep_row", u32),
        ("pad0", u32),
        ("acc_reset", u32 * STAT_COPIES),
        ("acc_timeout", u32 * STAT_COPIES),
        ("acc_ground", u32 * STAT_COPIES),
        ("acc_sums", (f32 * 8) * STAT_COPIES),
        ("ep_ring", (f32 * 8) * EP_RING),
    ]


class EnvState(C.Structure):
    _fields_ = [
        (k, vp)
        for k in (
            "pos",
            "quat",
            "vel",
            "ang",
            "commands",
            "last_actions",
            "pid_integral",
            "pid_prev_error",
            "episode_sums",
            "episode_length",
            "hover_counter",
        )
    ]


class EnvOut(C.Structure):
    _fields_ = [
        (k, vp)
        for k in (
            "obs",
            "rew",
            "reset",
            "time_outs",
            "base_lin_vel",
            "base_ang_vel",
            "rel_pos",
            "last_rel_pos",
            "base_euler",
            "crash",
            "rpm",
            "last_base_pos",
            "fixup_idx",
            "fixup_val",
        )
    ] + [("ep_row", u32), ("pad0", u32)]


class MpcConfig(C.Structure):
    _fields_ = [
        ("n", i32),
        ("m", i32),
        ("p", i32),
        ("L", i32),
        ("N", i32),
        ("alpha_reg_type", i32),
        ("max_iter", i32),
        ("gram_refresh", i32),
        ("alpha_sr_anchor", i32),
        ("reserved0", i32),
        ("Q", f64 * 8),
        ("R", f64 * 8),
        ("S", f64 * 8),
        ("lamb_alpha", f64),
        ("lamb_sigma", f64),
        ("lamb_alpha_s", f64),
        ("lamb_sigma_s", f64),
        ("U_lo", f64 * 8),
        ("U_hi", f64 * 8),
        ("Us_lo", f64 * 8),
        ("Us_hi", f64 * 8),
    ]


class MpcAux(C.Structure):
    _fields_ = [("u_ic", vp), ("y_ic", vp), ("v_sr", vp), ("sr_out", vp)]


i64 = C.c_int64
P = C.POINTER

# name -> (restype, argtypes); every symbol include/ddqc.h declares
SIGNATURES = {
    "ddqc_abi_version": (C.c_int, []),
    "ddqc_struct_size": (i64, [C.c_int]),
    "ddqc_env_step": (C.c_int, [P(EnvParams), P(EnvState), P(EnvOut), vp, vp, vp, i64, vp]),
    "ddqc_env_reset_idx": (C.c_int, [P(EnvParams), P(EnvState), P(EnvOut), vp, vp, i64, i64, vp]),
    "ddqc_env_get_state": (C.c_int, [P(EnvParams), P(EnvState), vp, i64, i64] + [vp] * 9 + [vp]),
    "ddqc_env_restore_state": (C.c_int, [P(EnvParams), P(EnvState), P(EnvOut), vp, vp, i64, i64] + [vp] * 7 + [vp]),
    "ddqc_env_update_target": (C.c_int, [P(EnvState), vp, i64, i64, vp, C.c_int, vp]),
    "ddqc_env_body_vel": (C.c_int, [P(EnvState), i64, vp, vp, vp]),
    "ddqc_env_flush_pid_reset": (C.c_int, [P(EnvState), vp, i64, vp]),
    "ddqc_pid_compute": (C.c_int, [vp, i64, i64, vp, i64, i64, vp, vp, P(f32), f32, vp, i64, i32, vp]),
    "ddqc_ctbr_compute": (
        C.c_int,
        [vp, i64, i64, vp, i64, i64, vp, i64, vp, vp, P(f32), P(f32), f32, f32, vp, i64, vp],
    ),
    "ddqc_tracking_compute": (
        C.c_int,
        [vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, vp, P(f32), f32, f32, f32, vp, i64, vp],
    ),
    "ddqc_ddmpc_workspace_bytes": (i64, [P(MpcConfig), i64]),
    "ddqc_ddmpc_gram_cache_bytes": (i64, [P(MpcConfig), i64]),
    "ddqc_ddmpc_solve": (C.c_int, [P(MpcConfig), vp] + [vp] * 11 + [i64, vp, vp]),
    "ddqc_ddmpc_solve_ex": (C.c_int, [P(MpcConfig), vp] + [vp] * 11 + [i64, vp, P(MpcAux), vp]),
    "ddqc_ddmpc_hankel_apply": (C.c_int, [P(MpcConfig), vp, vp, vp, vp, i64, vp]),
    "ddqc_ddmpc_recover_workspace_bytes": (i64, [P(MpcConfig), i64]),
    "ddqc_ddmpc_recover": (C.c_int, [P(MpcConfig), vp] + [vp] * 14 + [i64, vp]),
    "ddqc_ddmpc_refresh_data": (C.c_int, [P(MpcConfig), vp, vp, vp, vp, vp, C.c_double, vp, vp, i64, vp]),
    "ddqc_ddmpc_store": (C.c_int, [P(MpcConfig), vp, vp, vp, vp, i64, vp, vp]),
    "ddqc_ddmpc_action": (C.c_int, [P(MpcConfig), vp, C.c_int32, vp, vp, vp, i64, vp]),
    "ddqc_ddmpc_post_step": (C.c_int, [P(MpcConfig), vp, vp, vp, vp, i64, i64, vp, vp, C.c_double, vp, vp, vp, vp, C.c_int32, i64, vp, vp]),
    "ddqc_ddmpc_debug_phase_buffer": (C.c_int, [vp]),
    "ddqc_policy_mlp_forward": (C.c_int, [vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp]),
}

_STRUCTS = {0: EnvParams, 1: EnvCtl, 2: EnvState, 3: EnvOut, 4: MpcConfig, 5: MpcAux}

_lib = None


class DdqcError(RuntimeError):
    pass


def load():
    """Load libddqc.so and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DdqcError(
            f"{LIB_PATH} not found: the CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.ddqc_abi_version() != ABI_VERSION:
        raise DdqcError(f"{LIB_PATH} has ABI version {lib.ddqc_abi_version()}, this binding expects {ABI_VERSION}: rebuild")
    for which, cls in _STRUCTS.items():
        got = lib.ddqc_struct_size(which)
        if got != C.sizeof(cls):
            raise DdqcError(f"struct layout mismatch for {cls.__name__}: C {got} vs ctypes {C.sizeof(cls)}")
    _lib = lib
    return lib


_ERRORS = {-1: "null pointer", -2: "bad shape", -3: "misaligned pointer", -4: "unsupported configuration"}


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise DdqcError(f"{what}: {_ERRORS.get(rc, 'error')} ({rc})")
    raise DdqcError(f"{what}: CUDA error {rc}")


def require_cuda(device) -> "torch.device":  # noqa: F821
    import torch

    dev = torch.device(device)
    if dev.type != "cuda":
        raise DdqcError(
            f"device {dev} requested: this is the B200-native path and runs on CUDA devices only "
            "(no CPU fallback)."
        )
    if not torch.cuda.is_available():
        raise DdqcError("no CUDA device available; the B200-native path has no CPU fallback.")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class _NullGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL_GUARD = _NullGuard()


def device_guard(device):
    """Context manager that makes `device` the current CUDA device for a library call.

    The entry points launch on the *current* device (the stream passed in belongs to `device`), so an object built
    on `cuda:1` while `cuda:0` is current must switch for the duration of the call.  When `device` already is
    current - every single-GPU process, every torchrun rank after `torch.cuda.set_device` - this is a no-op that
    costs one `torch.cuda.current_device()` query."""
    import torch

    if torch.cuda.current_device() == device.index:
        return _NULL_GUARD
    return torch.cuda.device(device)


def stream_ptr(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
