#!/bin/bash
# env substep loop with the deferred velocity half vs the sequential loop; env parity + new evaluator tests;
# the 900-combination grid-search report on the pose-first physics
mkdir -p gpurun_out
{
for v in default seqv default seqv; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== $v"; python tools/env_launch_gap.py 2>&1 | tail -1
done
} > gpurun_out/r2_call24_env_variants.log 2>&1
unset DDQC_LIB
cat gpurun_out/r2_call24_env_variants.log
{
for v in default nopipe default nopipe; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  python tools/policy_variants.py 2>&1 | tail -1
done
} > gpurun_out/r2_call24_policy_variants.log 2>&1
unset DDQC_LIB
cat gpurun_out/r2_call24_policy_variants.log
timeout 1500 python -m pytest tests/test_env_parity_gpu.py tests/test_env_tile_kernel_gpu.py tests/test_ddmpc_eval_gpu.py tests/test_policy_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_call24_tests.log
cat gpurun_out/r2_call24_tests.log
timeout 600 python -m data_driven_quad_control_b200.data_driven_mpc.dd_mpc_param_grid_search --seed 0 --output-dir gpurun_out/grid_r2h > gpurun_out/r2_call24_grid.log 2>&1
tail -5 gpurun_out/r2_call24_grid.log; ls gpurun_out/grid_r2h
