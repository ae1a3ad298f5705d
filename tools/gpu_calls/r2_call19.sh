#!/bin/bash
# new DD-MPC modes (alpha_reg_type PREVIOUS, update_cost_threshold) + regression of the DD-MPC suites
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_ddmpc_parity_gpu.py tests/test_ddmpc_eval_gpu.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_call19_tests.log
cat gpurun_out/r2_call19_tests.log
