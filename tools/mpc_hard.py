"""Developer probe: the saved hard box-QP cases (gpurun_out/mpc_fail.npz) through the GPU solver vs the numpy mirror."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, functools
import ddmpc_oracle as mo, ddmpc_condensed as dc
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import BatchedNonlinearDataDrivenMPC
d = np.load(os.path.join(ROOT, "tests", "golden", "ddmpc_hard_cases.npz"))
n = d["u"].shape[0]
prm = mo.default_params()
kw = mo.ctor_kwargs(prm, d["u"][0], d["y"][0], np.array([1, 1, 2.5]))
kw["u"], kw["y"], kw["y_r"] = d["u"], d["y"], np.tile(np.array([[1, 1, 2.5]]), (n, 1))
ctrl = BatchedNonlinearDataDrivenMPC(**kw, device="cuda", solve_on_init=False)
ctrl._act_state = None
ctrl.update_and_solve_data_driven_mpc(); torch.cuda.synchronize()
print("status", ctrl.status.cpu().tolist(), "passes", ctrl.pivoting_passes.cpu().tolist(), "ipm", ctrl.ipm_iterations.cpu().tolist())
err = np.abs(ctrl.optimal_u.cpu().numpy() - d["u_opt"]).max(axis=(1, 2))
print("max |u - u_ref|", err.max(), err)
