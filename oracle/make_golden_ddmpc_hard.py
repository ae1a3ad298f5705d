"""TEST INFRASTRUCTURE -- regenerates tests/golden/ddmpc_hard_cases.npz.

The cases are default-size (n=6, L=35, N=350) closed-loop windows collected on the GPU env during
round 1 on which block principal pivoting from a cold start cycles (54-84 of the 93 box-QP
variables sit on their bounds); `tools/mpc_fail.py` dumps such windows into gpurun_out/mpc_fail.npz.
This script attaches, for every case, the solution of the LITERAL 693-variable QP by the
interior-point oracle (`ddmpc_oracle.NonlinearDDMPCOracle`, KKT-certified) so that the fixture is
anchored on the oracle and not on any condensed algorithm.

  python oracle/make_golden_ddmpc_hard.py [gpurun_out/mpc_fail.npz]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ddmpc_oracle as mo  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "ddmpc_hard_cases.npz")


def main():
    prm = mo.default_params()
    y_r = np.array([1.0, 1.0, 2.5])
    if len(sys.argv) > 1:
        d = np.load(sys.argv[1])
        sel = [3, 9, 12, 16, 18, 28, 36, 39]
        U = np.array([d[f"u_{i}"] for i in sel])
        Y = np.array([d[f"y_{i}"] for i in sel])
        UP = np.array([d[f"us_prev_{i}"] for i in sel])
        YP = np.array([d[f"ys_prev_{i}"] for i in sel])
    else:  # refresh the oracle solutions of the existing fixture
        d = np.load(OUT)
        U, Y, UP, YP = d["u"], d["y"], d["us_prev"], d["ys_prev"]
    UO, US, YS, KKT = [], [], [], []
    for b in range(U.shape[0]):
        orc = mo.NonlinearDDMPCOracle(**mo.ctor_kwargs(prm, U[b], Y[b], y_r), solve_on_init=False)
        orc.us_prev, orc.ys_prev = UP[b].copy(), YP[b].copy()
        orc.update_and_solve_data_driven_mpc()
        k = mo.kkt_residuals(*orc.last["qp"], orc.last["z"], *orc.last["mult"])
        UO.append(orc.optimal_u)
        US.append(orc.us_prev)
        YS.append(orc.ys_prev)
        KKT.append([k["stationarity"], k["equality"], k["inequality"], k["min_active_multiplier"]])
        print(b, KKT[-1], "active bounds:", int((np.abs(orc.optimal_u - prm.U[:, 0]) < 1e-7).sum() + (np.abs(orc.optimal_u - prm.U[:, 1]) < 1e-7).sum()))
    np.savez_compressed(OUT, u=U, y=Y, us_prev=UP, ys_prev=YP, u_opt=np.array(UO), us=np.array(US), ys=np.array(YS),
                        kkt=np.array(KKT), y_r=y_r)


if __name__ == "__main__":
    main()
