"""Developer probe: dump DD-MPC problems whose pivoting hit the iteration cap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from data_driven_quad_control_b200.data_driven_mpc import nonlinear_data_driven_mpc_controller as mod
orig = mod.BatchedNonlinearDataDrivenMPC.update_and_solve_data_driven_mpc
dump = []
def patched(self):
    u, y, usp, ysp, hp = self.u.clone(), self.y.clone(), self.us.clone(), self.ys.clone(), self._have_prev.clone()
    orig(self)
    bad = (self.status != 0).nonzero().flatten()
    slow = (self.iters > 30).nonzero().flatten()
    for b in torch.unique(torch.cat([bad, slow])).tolist()[:6]:
        dump.append(dict(u=u[b].cpu().numpy(), y=y[b].cpu().numpy(), us_prev=usp[b].cpu().numpy(), ys_prev=ysp[b].cpu().numpy(),
                         have_prev=int(hp[b]), status=int(self.status[b]), iters=int(self.iters[b]), u_opt=self.optimal_u[b].cpu().numpy()))
mod.BatchedNonlinearDataDrivenMPC.update_and_solve_data_driven_mpc = patched
leg = bench.run_ddmpc_leg(torch.device("cuda"), 4096, 12)
print({k: leg[k] for k in ("value", "pivoting_passes_mean", "pivoting_passes_max", "failed_controllers")}, "dumped", len(dump))
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "mpc_fail.npz"), **{f"{k}_{i}": v for i, d in enumerate(dump[:40]) for k, v in d.items()})
