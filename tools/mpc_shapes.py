"""Developer probe: GPU solver vs literal oracle over several (n, L, N) shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import ddmpc_oracle as mo
from test_ddmpc_parity_gpu import _synthetic_windows, _batched
for (n, L, N) in [(2, 5, 80), (2, 6, 80), (2, 7, 90), (3, 6, 100), (3, 9, 130), (4, 10, 160)]:
    prm = mo.default_params(); prm.n, prm.L, prm.N = n, L, N; prm.alpha_reg_type = 0
    u, y = _synthetic_windows(prm, 3, 2)
    y_r = np.tile(np.array([[0.3, -0.2, 0.5]]), (2, 1))
    try:
        gpu = _batched(prm, u, y, y_r, solve_on_init=False)
    except Exception as e:
        print((n, L, N), "ctor:", e); continue
    errs = []
    orcs = [mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, u[b], y[b], y_r[b]), solve_on_init=False) for b in range(2)]
    for t in range(2):
        gpu.update_and_solve_data_driven_mpc()
        for b, orc in enumerate(orcs):
            orc.update_and_solve_data_driven_mpc()
            errs.append(np.abs(gpu.optimal_u[b].cpu().numpy() - orc.optimal_u).max())
            u0 = orc.optimal_u[0]
            orc.store_input_output_measurement(u0, y[b, -1])
        gpu.store_input_output_measurement(torch.from_numpy(np.stack([o.u[-1] for o in orcs])).cuda(), torch.from_numpy(y[:, -1]).cuda())
    print((n, L, N), "status", gpu.status.cpu().tolist(), "max err", max(errs))
