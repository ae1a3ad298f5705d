#!/bin/bash
# the optional DD-MPC modes only (after the tolerance note), then the host-logic + evaluation suites
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_ddmpc_parity_gpu.py -q -m gpu -k "previous or threshold or unsupported or numpy_api" 2>&1 | tail -30 > gpurun_out/r2_call20_tests.log
cat gpurun_out/r2_call20_tests.log
