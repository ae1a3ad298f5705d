#!/bin/bash
# round-2 GPU call 14: branch-free sqrt in the env kernels: parity + A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_reference_behaviour_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2c14_tests.log
{
for v in default sqrtold default sqrtold; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== $v"; timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
done
unset DDQC_LIB
echo "== 4096"; timeout 120 python tools/env_launch_gap.py 4096 2>&1 | tail -1
} > gpurun_out/r2c14_timing.log 2>&1
cat gpurun_out/r2c14_tests.log gpurun_out/r2c14_timing.log
