"""Developer probe: DD-MPC solver phase timing at default sizes."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from data_driven_quad_control_b200 import _lib
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import BatchedNonlinearDataDrivenMPC
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
g = np.load(os.path.join(ROOT, "tests", "golden", "ddmpc_default.npz"))
rng = np.random.default_rng(0)
u = np.repeat(g["u"][None], B, 0) + 1e-3 * rng.standard_normal((B, 350, 3))
y = np.repeat(g["y"][None], B, 0) + 1e-4 * rng.standard_normal((B, 350, 3))
P = bench.MPC_DEFAULT; ell = 42
ctrl = BatchedNonlinearDataDrivenMPC(n=6, m=3, p=3, u=u, y=y, L=35, Q=np.kron(np.eye(ell), np.diag(P["Q"])), R=np.kron(np.eye(ell), np.diag(P["R"])), S=np.diag(P["S"]), y_r=np.array(P["y_r"]), lamb_alpha=50.0, lamb_sigma=1e9, U=P["U"], Us=P["Us"], alpha_reg_type=0, lamb_alpha_s=1.0, lamb_sigma_s=1e7, solve_on_init=False,
                                     gram_cache=os.environ.get("GRAM_CACHE", "1") == "1")
lib = _lib.load()
clk = torch.zeros(48, dtype=torch.int64, device="cuda")
ctrl.update_and_solve_data_driven_mpc(); torch.cuda.synchronize()
# one closed-loop step: the timed solve sees a window that moved by one sample (Gram cache: rank-2 update path)
ctrl.store_input_output_measurement(ctrl.get_optimal_control_input_at_step(0).clone(), ctrl.y[:, -1].clone() + 1e-3)
torch.cuda.synchronize()
lib.ddqc_ddmpc_debug_phase_buffer(clk.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctrl.update_and_solve_data_driven_mpc(); e1.record(); torch.cuda.synchronize()
lib.ddqc_ddmpc_debug_phase_buffer(None)
ms = e0.elapsed_time(e1)
names = ["gram", "A_factor", "A_update", "B_factor", "B_backsub", "C_factor_schur", "sweep", "qp_assemble", "pivoting", "ipm_polish", "out", "-"]
call = clk.cpu().numpy().astype(float); c = call[:12]
if call[24:].sum() > 0: print("sub-phases kcycles/solve:", (call[24:] / B / 1e3).round(1))
if call[12:24].sum() > 0: print("factor stages kcycles/solve  w0:", (call[12:18] / B / 1e3).round(1), " w1:", (call[18:24] / B / 1e3).round(1))
print(json.dumps({"B": B, "ms": ms, "solves_per_s": B / ms * 1e3, "us_per_solve_per_cta": ms * 1e3 * min(B, 148) / B,
                  "phase_pct": {n: round(100 * v / c.sum(), 1) for n, v in zip(names, c)},
                  "phase_kcycles_per_solve": {n: round(v / B / 1e3, 1) for n, v in zip(names, c)},
                  "passes_mean": float(ctrl.pivoting_passes.float().mean()), "ipm_frac": float((ctrl.ipm_iterations > 0).float().mean()), "status_bad": int((ctrl.status != 0).sum())}))
