"""Developer probe: cost of the optional DD-MPC modes at 4096 default-size controllers (same window for every controller):
ms per batched update_and_solve for alpha_reg_type ZERO (no recovery), PREVIOUS (v_sr + solve + recovery) and ZERO with
update_cost_threshold (refresh + solve without Gram cache + recovery)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import ddmpc_oracle as mo
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import BatchedNonlinearDataDrivenMPC
g = np.load(os.path.join(ROOT, "tests", "golden", "ddmpc_default.npz"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
out = {"lib": os.environ.get("DDQC_LIB", "default"), "B": B}
for name, art, kw in (("zero", 2, {}), ("previous", 1, {}), ("zero_threshold", 2, dict(update_cost_threshold=1e-3))):
    prm = mo.default_params(); prm.alpha_reg_type = art
    k = mo.ctor_kwargs(prm, g["u"], g["y"], g["y_r"])
    k.update(u=np.tile(g["u"][None], (B, 1, 1)), y=np.tile(g["y"][None], (B, 1, 1)), y_r=np.tile(g["y_r"][None], (B, 1)))
    c = BatchedNonlinearDataDrivenMPC(**{**k, **kw}, device="cuda", solve_on_init=False)
    ms = []
    for t in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c.update_and_solve_data_driven_mpc(); e1.record()
        c.store_input_output_measurement(c.get_optimal_control_input_at_step(0), torch.from_numpy(g["y_next"][min(t, 2)][None]).cuda().repeat(B, 1))
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    out[name] = {"ms": round(sum(ms[1:]) / 4, 3), "bad": int((c.status != 0).sum())}
print(json.dumps(out))
