import sys, os
ROOT=os.path.dirname(os.path.abspath(__file__))
sys.path[:0]=[ROOT, os.path.join(ROOT,"oracle"), os.path.join(ROOT,"tests")]
import numpy as np, torch
import ddmpc_oracle as mo, ddmpc_condensed as mc
from test_ddmpc_parity_gpu import _small_params, _synthetic_windows, _batched
prm=_small_params(); prm.alpha_reg_type=0
B=6
u,y=_synthetic_windows(prm,3,B)
y_r = np.tile(np.array([[0.3, -0.2, 0.5]]), (B, 1)) + 0.1 * np.arange(B)[:, None]
gpu=_batched(prm,u,y,y_r,solve_on_init=False)
us_prev=[None]*B; ys_prev=[None]*B
rng=np.random.default_rng(0)
for t in range(3):
    gpu.update_and_solve_data_driven_mpc()
    print("t",t,"status",gpu.status.tolist(),"iters",gpu.iters.tolist())
    uw=gpu.u.cpu().numpy(); yw=gpu.y.cpu().numpy()
    for b in range(B):
        try:
            res=mc.solve_condensed(prm,uw[b],yw[b],y_r[b],us_prev[b],ys_prev[b])
            print(b, "max|du|", np.abs(gpu.optimal_u[b].cpu().numpy()-res["u_opt"]).max(), "d us", np.abs(gpu.us[b].cpu().numpy()-res["us"]).max(), "np iters",res["iters"], "|v_sr|", np.abs(res["v_sr"]).max())
            us_prev[b],ys_prev[b]=res["us"],res["ys"]
        except Exception as e:
            print(b,"numpy failed",e)
    u0=gpu.get_optimal_control_input_at_step(0).cpu().numpy()
    y_new=rng.standard_normal((B,3))*0.05
    gpu.store_input_output_measurement(u0, yw[:,-1]+y_new)
