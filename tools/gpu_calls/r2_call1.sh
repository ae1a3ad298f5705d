#!/bin/bash
# round-2 GPU call 1: reference-stack probe, env parity on the new two-launch step, env kernel variants
mkdir -p gpurun_out
tools/probe_reference_stack.sh > /dev/null 2>&1
python -m pytest tests/test_env_parity_gpu.py tests/test_reference_behaviour_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2c1_env_tests.log
DDQC_LIB=tools/variants/lib_fin1.so python -m pytest tests/test_env_parity_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2c1_env_tests_fin1.log
{
for v in default fin0 fin1 fin2_t256 fin2_t64; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== $v"
  python tools/env_perf.py 2>&1 | tail -1
  python tools/env_launch_gap.py 2>&1 | tail -1
  python tools/env_perf.py 131072 2>&1 | tail -1
  python tools/env_perf.py 4096 2>&1 | tail -1
done
} > gpurun_out/r2c1_env_variants.log 2>&1
unset DDQC_LIB
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c1_gpu_tests.log
cat gpurun_out/r2c1_env_tests.log gpurun_out/r2c1_env_tests_fin1.log gpurun_out/r2c1_env_variants.log gpurun_out/r2c1_gpu_tests.log
