"""Batched nonlinear data-driven MPC controllers on the `ddqc_ddmpc_solve` CUDA kernel.

The reference drives a third-party `direct_data_driven_mpc.NonlinearDataDrivenMPCController`
(one CPU process per controller, cvxpy + OSQP per step).  This module keeps that controller's
constructor keywords and step API -- as used by the reference at
`data_driven_mpc/utilities/param_grid_search/controller_creation.py:91-112` (ctor),
`.../controller_evaluation.py:361` (`update_and_solve_data_driven_mpc`), `:389-392`
(`get_optimal_control_input_at_step`), `:416-422` (`get_du_value_at_step`,
`store_input_output_measurement`) and `comparison/utilities/controller_workers/dd_mpc_worker.py:78`
(`set_output_setpoint`) -- and adds a batched front: `BatchedNonlinearDataDrivenMPC` holds B
controllers (leading dimension B on every tensor) and solves them in one launch.

`NonlinearDataDrivenMPCController` is the single-controller, numpy-in / numpy-out drop-in.

Supported configuration: `ext_out_incr_in=False` (what every shipped reference config uses), all three
`alpha_reg_type`s, `update_cost_threshold`, `n_mpc_step`, block-diagonal Q / R with identical diagonal blocks
(what `controller_creation.py:73-88` builds).  Anything else raises NotImplementedError / ValueError instead of
silently solving a different problem.

The two modes that need the optimal alpha - which the condensed solver never forms - run a dual-recovery
kernel after every solve (`ddqc_ddmpc_recover`, csrc/ddqc_ddmpc_recover.cu):

* `alpha_reg_type` PREVIOUS (1): alpha_sr of a solve = the optimal alpha of the previous solve (zeros on the
  first); `self.alpha_prev` (B, C) holds it, `ddqc_ddmpc_hankel_apply` turns it into v_sr = H alpha_sr on the
  current data before the solve.
* `update_cost_threshold`: the measurement arrays `self.u / self.y` always roll (the initial condition is always the
  latest n samples); the Hankel data of a solve is a snapshot of them (`self.u_data / self.y_data`), refreshed
  before the solve unless the previous solve's tracking cost (`self.tracking_cost`) was <= the threshold
  (configs/data_driven_mpc/dd_mpc_controller_params.yaml:85-90; semantics restated in oracle/ddmpc_oracle.py).
"""

from __future__ import annotations

import os

import ctypes as C
from enum import Enum

import numpy as np
import torch

from data_driven_quad_control_b200 import _lib


ALPHA_SR_ANCHORS = {"optimal_reachable_equilibrium": 0, "previous_equilibrium": 1}


class AlphaRegType(Enum):
    APPROXIMATED = 0  # regularise w.r.t. an approximation of alpha_Lin^sr(D_t)
    PREVIOUS = 1  # w.r.t. the previous optimal alpha
    ZERO = 2  # w.r.t. zero


def _block_diag_weights(M, blocks: int, dim: int, name: str) -> np.ndarray:
    """Diagonal of the repeated block of kron(I_blocks, diag(w)); validates the structure."""
    M = np.asarray(M, dtype=np.float64)
    if M.shape != (blocks * dim, blocks * dim):
        raise ValueError(f"{name} must have shape {(blocks * dim, blocks * dim)}, got {M.shape}")
    w = np.diag(M)[:dim].copy()
    if not np.array_equal(M, np.kron(np.eye(blocks), np.diag(w))):
        raise NotImplementedError(f"{name} must be block diagonal with identical diagonal blocks: kron(I, diag(w))")
    if (w <= 0).any():
        raise ValueError(f"{name} weights must be positive")
    return w


class BatchedNonlinearDataDrivenMPC:
    def __init__(
        self,
        n: int,
        m: int,
        p: int,
        u,
        y,
        L: int,
        Q,
        R,
        S,
        y_r,
        lamb_alpha: float,
        lamb_sigma: float,
        U,
        Us,
        alpha_reg_type: AlphaRegType | int = AlphaRegType.ZERO,
        lamb_alpha_s: float | None = None,
        lamb_sigma_s: float | None = None,
        ext_out_incr_in: bool = False,
        update_cost_threshold: float | None = None,
        n_mpc_step: int = 1,
        device: torch.device | str = "cuda",
        solve_on_init: bool = True,
        max_iter: int = 12,
        warm_start: bool = True,
        gram_cache: bool | None = None,
        gram_refresh: int = 128,
        alpha_sr_anchor: str = "optimal_reachable_equilibrium",
    ):
        """`alpha_sr_anchor` (extension, alpha_reg_type APPROXIMATED only) selects which equilibrium the approximation
        of alpha_Lin^sr(D_t) belongs to - the third-party source is not available and the published scheme leaves it
        open: "optimal_reachable_equilibrium" (default; the equilibrium closest to y_r) or
        "previous_equilibrium" (the optimal (u_s, y_s) of the previous solve; alpha_sr = 0 on the first solve)."""
        if alpha_sr_anchor not in ALPHA_SR_ANCHORS:
            raise ValueError(f"alpha_sr_anchor must be one of {sorted(ALPHA_SR_ANCHORS)}")
        if ext_out_incr_in:
            raise NotImplementedError("ext_out_incr_in=True (extended output / input increments) is not supported")
        art = AlphaRegType(alpha_reg_type) if not isinstance(alpha_reg_type, AlphaRegType) else alpha_reg_type
        # null or 0 = "input-output data is always updated online" (dd_mpc_controller_params.yaml:89-90)
        self.update_cost_threshold = float(update_cost_threshold) if update_cost_threshold else None
        if art == AlphaRegType.APPROXIMATED and (lamb_alpha_s is None or lamb_sigma_s is None):
            raise ValueError("lamb_alpha_s and lamb_sigma_s are required for alpha_reg_type APPROXIMATED")
        self._lib = _lib.load()
        self.device = _lib.require_cuda(device)
        dev = self.device
        f64 = dict(dtype=torch.float64, device=dev)

        u_t = torch.as_tensor(np.asarray(u) if not torch.is_tensor(u) else u).to(**f64)
        y_t = torch.as_tensor(np.asarray(y) if not torch.is_tensor(y) else y).to(**f64)
        if u_t.ndim != 3 or y_t.ndim != 3 or u_t.shape[:2] != y_t.shape[:2] or u_t.shape[2] != m or y_t.shape[2] != p:
            raise ValueError("u must have shape (B, N, m) and y (B, N, p)")
        self.B, self.N = int(u_t.shape[0]), int(u_t.shape[1])
        self.n, self.m, self.p, self.L = int(n), int(m), int(p), int(L)
        self.n_mpc_step = int(n_mpc_step)
        if not 1 <= self.n_mpc_step <= self.L + 1:
            raise ValueError("n_mpc_step out of range")
        ell = self.L + self.n + 1
        if self.N - ell + 1 < (m + p) * ell + 1:
            # fewer Hankel columns than rows: the data cannot be persistently exciting of the needed order
            raise ValueError(f"N = {self.N} is too short for L = {L}, n = {n}: need N >= {(m + p + 1) * ell}")
        self.alpha_reg_type = art
        self.u = u_t.contiguous().clone()
        self.y = y_t.contiguous().clone()

        cfg = _lib.MpcConfig()
        cfg.n, cfg.m, cfg.p, cfg.L, cfg.N = self.n, self.m, self.p, self.L, self.N
        cfg.alpha_reg_type = art.value
        cfg.max_iter = int(max_iter)
        cfg.gram_refresh = int(gram_refresh)
        cfg.alpha_sr_anchor = ALPHA_SR_ANCHORS[alpha_sr_anchor]
        self.alpha_sr_anchor = alpha_sr_anchor
        qw = _block_diag_weights(Q, ell, p, "Q")
        rw = _block_diag_weights(R, ell, m, "R")
        sw = _block_diag_weights(S, 1, p, "S")
        U = np.asarray(U, dtype=np.float64).reshape(m, 2)
        Us = np.asarray(Us, dtype=np.float64).reshape(m, 2)
        for i in range(p):
            cfg.Q[i], cfg.S[i] = qw[i], sw[i]
        for i in range(m):
            cfg.R[i] = rw[i]
            cfg.U_lo[i], cfg.U_hi[i] = U[i]
            cfg.Us_lo[i], cfg.Us_hi[i] = Us[i]
        # regularisation weights: scalars (one value for the batch) or (B,) arrays (one per controller)
        lam = [lamb_alpha, lamb_sigma, 1.0 if lamb_alpha_s is None else lamb_alpha_s, 1.0 if lamb_sigma_s is None else lamb_sigma_s]
        per_ctrl = any(np.ndim(v) > 0 or torch.is_tensor(v) for v in lam)
        if per_ctrl:
            cols = [torch.as_tensor(np.asarray(v.cpu()) if torch.is_tensor(v) else np.asarray(v), dtype=torch.float64).reshape(-1) for v in lam]
            cols = [c.expand(self.B) if c.numel() == 1 else c for c in cols]
            if any(c.numel() != self.B for c in cols):
                raise ValueError("per-controller regularisation weights must have one entry per controller")
            self._lamb = torch.stack(cols, dim=1).to(dev).contiguous()
            if not bool((self._lamb > 0).all()):
                raise ValueError("regularisation weights must be positive")
            lam = [float(c[0]) for c in cols]
        else:
            self._lamb = None
        cfg.lamb_alpha, cfg.lamb_sigma, cfg.lamb_alpha_s, cfg.lamb_sigma_s = (float(v) for v in lam)
        self._cfg = cfg
        ws_bytes = self._lib.ddqc_ddmpc_workspace_bytes(C.byref(cfg), self.B)
        if ws_bytes < 0:
            raise ValueError("unsupported problem size for the GPU solver (see ddqc_ddmpc.cu limits)")
        self._ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        # Per-controller Gram cache (ddqc.h): G persists in HBM between solves and follows the one-sample window shift
        # by a rank-2 correction.  `gram_cache=None` enables it when it fits in a quarter of the free device memory.
        gc_bytes = self._lib.ddqc_ddmpc_gram_cache_bytes(C.byref(cfg), self.B)
        if self.update_cost_threshold is not None:
            if gram_cache:
                raise ValueError("gram_cache follows one-sample window shifts; it cannot be combined with update_cost_threshold")
            gram_cache = False  # a frozen / re-snapshotted data window is not a one-sample shift
        if gram_cache is None:
            gram_cache = gc_bytes <= torch.cuda.mem_get_info(dev)[0] // 4 and os.environ.get("DDQC_GRAM_CACHE", "1") != "0"
        self._gram_cache = torch.zeros((gc_bytes,), dtype=torch.uint8, device=dev) if gram_cache else None

        self.y_r = torch.zeros((self.B, p), **f64)
        self.set_output_setpoint(y_r)
        self.us = torch.zeros((self.B, m), **f64)  # last optimal equilibrium input
        self.ys = torch.zeros((self.B, p), **f64)
        self._have_prev = torch.zeros((self.B,), dtype=torch.int32, device=dev)
        self.optimal_u = torch.zeros((self.B, self.L + 1, m), **f64)
        self.status = torch.zeros((self.B,), dtype=torch.int32, device=dev)
        self.iters = torch.zeros((self.B,), dtype=torch.int32, device=dev)
        # active set of the condensed box QP, carried between solves as the warm start
        self.nx = (self.L - self.n) * m + m + p
        self._act_state = torch.zeros((self.B, self.nx), dtype=torch.int8, device=dev) if warm_start else None
        # ---- optional modes (module docstring) --------------------------------------------------------------
        thr, nsig, C_ = self.update_cost_threshold, m + p, self.N - ell + 1
        self._needs_recover = art == AlphaRegType.PREVIOUS or thr is not None
        self.alpha_prev = torch.zeros((self.B, C_), **f64) if art == AlphaRegType.PREVIOUS else None
        self._v_sr = torch.zeros((self.B, nsig * ell + 1), **f64) if art == AlphaRegType.PREVIOUS else None
        self._sr = torch.zeros((self.B, nsig + p * ell), **f64) if (thr is not None and art == AlphaRegType.APPROXIMATED) else None
        self.tracking_cost = torch.full((self.B,), float("inf"), **f64) if thr is not None else None
        self.data_updated = torch.ones((self.B,), dtype=torch.int32, device=dev) if thr is not None else None
        self.u_data = self.u.clone() if thr is not None else self.u  # Hankel data of the solves
        self.y_data = self.y.clone() if thr is not None else self.y
        if self._needs_recover:
            rb = self._lib.ddqc_ddmpc_recover_workspace_bytes(C.byref(cfg), self.B)
            if rb < 0:
                raise ValueError("unsupported problem size for the dual-recovery kernel (ddqc_ddmpc_recover.cu)")
            self._rec_ws = torch.empty((rb,), dtype=torch.uint8, device=dev)
        self._aux = _lib.MpcAux()
        if thr is not None:
            self._aux.u_ic, self._aux.y_ic = self.u.data_ptr(), self.y.data_ptr()
        if self._v_sr is not None:
            self._aux.v_sr = self._v_sr.data_ptr()
        if self._sr is not None:
            self._aux.sr_out = self._sr.data_ptr()
        if solve_on_init:
            self.update_and_solve_data_driven_mpc()

    # ------------------------------------------------------------------ API
    def set_output_setpoint(self, y_r) -> None:
        t = torch.as_tensor(np.asarray(y_r) if not torch.is_tensor(y_r) else y_r).to(self.y_r)
        if t.numel() == self.p:
            t = t.reshape(1, self.p).expand(self.B, self.p)
        self.y_r.copy_(t.reshape(self.B, self.p))

    def update_and_solve_data_driven_mpc(self, skip: torch.Tensor | None = None) -> None:
        """One solve of every controller of the batch on its current (u, y) window.  `skip` ((B,) bool, optional):
        controllers left out of this solve - their outputs and status keep their last values (used by the batched
        evaluator for runs the reference would have aborted)."""
        have_prev = self._have_prev if skip is None else torch.where(skip, -1, self._have_prev).to(torch.int32)
        cfg, st, thr = C.byref(self._cfg), _lib.stream_ptr(self.device), self.update_cost_threshold
        lamb = self._lamb.data_ptr() if self._lamb is not None else None
        with _lib.device_guard(self.device):
            if thr is not None:  # refresh the Hankel data of the controllers whose last tracking cost exceeds the threshold
                _lib.check(self._lib.ddqc_ddmpc_refresh_data(
                    cfg, self.u.data_ptr(), self.y.data_ptr(), self.u_data.data_ptr(), self.y_data.data_ptr(),
                    self.tracking_cost.data_ptr(), thr, have_prev.data_ptr(), self.data_updated.data_ptr(), self.B, st,
                ), "ddqc_ddmpc_refresh_data")  # fmt: skip
            if self.alpha_prev is not None:  # v_sr = H alpha_prev on the data of this solve
                _lib.check(self._lib.ddqc_ddmpc_hankel_apply(
                    cfg, self.u_data.data_ptr(), self.y_data.data_ptr(), self.alpha_prev.data_ptr(), self._v_sr.data_ptr(), self.B, st,
                ), "ddqc_ddmpc_hankel_apply")  # fmt: skip
            rc = self._lib.ddqc_ddmpc_solve_ex(
                cfg, self._ws.data_ptr(), self.u_data.data_ptr(), self.y_data.data_ptr(), self.y_r.data_ptr(),
                have_prev.data_ptr(), lamb,
                self._act_state.data_ptr() if self._act_state is not None else None, self.optimal_u.data_ptr(),
                self.us.data_ptr(), self.ys.data_ptr(), self.status.data_ptr(), self.iters.data_ptr(), self.B,
                self.gram_cache_ptr, C.byref(self._aux), st,
            )  # fmt: skip
            _lib.check(rc, "ddqc_ddmpc_solve")
            if self._needs_recover:
                _lib.check(self._lib.ddqc_ddmpc_recover(
                    cfg, self._rec_ws.data_ptr(), self.u_data.data_ptr(), self.y_data.data_ptr(),
                    self._aux.u_ic, self._aux.y_ic, self.y_r.data_ptr(), have_prev.data_ptr(), lamb,
                    self.optimal_u.data_ptr(), self.us.data_ptr(), self.ys.data_ptr(), self._aux.v_sr, self._aux.sr_out,
                    self.alpha_prev.data_ptr() if self.alpha_prev is not None else None,
                    self.tracking_cost.data_ptr() if self.tracking_cost is not None else None, self.B, st,
                ), "ddqc_ddmpc_recover")  # fmt: skip
        if skip is None:
            self._have_prev.fill_(1)
        else:  # a skipped controller keeps the history it had (it may be solved again later)
            self._have_prev.masked_fill_(~skip, 1)

    @property
    def gram_cache_ptr(self):
        """Device pointer of the Gram cache (None when it is off), for the C-ABI calls that shift the windows."""
        return self._gram_cache.data_ptr() if self._gram_cache is not None else None

    def invalidate_gram_cache(self) -> None:
        """Call after writing `self.u` / `self.y` directly: the next solve regenerates G from the windows."""
        if self._gram_cache is not None:
            self._gram_cache.zero_()

    @property
    def pivoting_passes(self) -> torch.Tensor:
        """Pivoting passes of the last solve (per controller)."""
        return self.iters & 0xFFFF

    @property
    def ipm_iterations(self) -> torch.Tensor:
        """Interior-point iterations of the last solve (0 unless the pivoting fell back)."""
        return self.iters >> 16

    def get_optimal_control_input_at_step(self, n_step: int = 0) -> torch.Tensor:
        if not 0 <= n_step < self.n_mpc_step:
            raise ValueError("The specified step `n_step` must be within the range [0, n_mpc_step - 1].")
        return self.optimal_u[:, n_step, :]

    def get_du_value_at_step(self, n_step: int = 0):
        return None  # only meaningful with ext_out_incr_in=True

    def store_input_output_measurement(self, u_current, y_current, du_current=None) -> None:
        u_c = torch.as_tensor(np.asarray(u_current) if not torch.is_tensor(u_current) else u_current)
        y_c = torch.as_tensor(np.asarray(y_current) if not torch.is_tensor(y_current) else y_current)
        u_c = u_c.to(device=self.device, dtype=torch.float64).reshape(self.B, self.m).contiguous()
        y_c = y_c.to(device=self.device, dtype=torch.float64).reshape(self.B, self.p).contiguous()
        with _lib.device_guard(self.device):
            rc = self._lib.ddqc_ddmpc_store(
                C.byref(self._cfg), self.u.data_ptr(), self.y.data_ptr(), u_c.data_ptr(), y_c.data_ptr(), self.B,
                self.gram_cache_ptr, _lib.stream_ptr(self.device),
            )
        _lib.check(rc, "ddqc_ddmpc_store")

    def check_status(self) -> None:
        """Raises if any controller of the batch failed (synchronises).  Mirrors the per-evaluation
        FAILURE of the reference (controller_evaluation.py:178-215)."""
        st = self.status.cpu()
        if (st != 0).any():
            bad = st.nonzero().flatten().tolist()
            raise RuntimeError(f"DD-MPC solve failed for controllers {bad[:8]} (status {st[bad[0]].item()})")


class NonlinearDataDrivenMPCController:
    """Single-controller drop-in with the numpy interface of the third-party controller."""

    def __init__(self, n, m, p, u, y, L, Q, R, S, y_r, lamb_alpha, lamb_sigma, U, Us, alpha_reg_type=AlphaRegType.ZERO,
                 lamb_alpha_s=None, lamb_sigma_s=None, ext_out_incr_in=False, update_cost_threshold=None,
                 n_mpc_step=1, device="cuda", solve_on_init=True, alpha_sr_anchor="optimal_reachable_equilibrium"):  # fmt: skip
        u, y = np.asarray(u, dtype=np.float64), np.asarray(y, dtype=np.float64)
        self._b = BatchedNonlinearDataDrivenMPC(
            n, m, p, u[None], y[None], L, Q, R, S, np.asarray(y_r, dtype=np.float64).reshape(1, -1), lamb_alpha,
            lamb_sigma, U, Us, alpha_reg_type, lamb_alpha_s, lamb_sigma_s, ext_out_incr_in, update_cost_threshold,
            n_mpc_step, device, solve_on_init, alpha_sr_anchor=alpha_sr_anchor,
        )  # fmt: skip
        self.n, self.m, self.p, self.L, self.N = n, m, p, L, u.shape[0]
        self.n_mpc_step = n_mpc_step
        self.y_r = np.asarray(y_r, dtype=np.float64).reshape(-1, 1)
        if solve_on_init:
            self._b.check_status()

    def set_output_setpoint(self, y_r) -> None:
        self.y_r = np.asarray(y_r, dtype=np.float64).reshape(-1, 1)
        self._b.set_output_setpoint(self.y_r.reshape(1, -1))

    def update_and_solve_data_driven_mpc(self) -> None:
        self._b.update_and_solve_data_driven_mpc()
        self._b.check_status()

    def get_optimal_control_input_at_step(self, n_step: int = 0) -> np.ndarray:
        return self._b.get_optimal_control_input_at_step(n_step)[0].cpu().numpy()

    def get_du_value_at_step(self, n_step: int = 0):
        return None

    def store_input_output_measurement(self, u_current, y_current, du_current=None) -> None:
        self._b.store_input_output_measurement(
            np.asarray(u_current, dtype=np.float64).reshape(1, -1), np.asarray(y_current, dtype=np.float64).reshape(1, -1)
        )
