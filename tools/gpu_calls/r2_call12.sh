#!/bin/bash
# round-2 GPU call 12: pair kernel variants (register budget), negation sharing
mkdir -p gpurun_out
export DDQC_ENV_PAIR_TILE_MIN_N=100000000000
timeout 600 python -m pytest tests/test_env_tile_kernel_gpu.py -x -q -m gpu -k "pair and not pair_tile" 2>&1 | tail -5 > gpurun_out/r2c12_tests.log
{
for v in default pb5 pb6 pt256b2; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== pair $v"; timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
done
} > gpurun_out/r2c12_timing.log 2>&1
cat gpurun_out/r2c12_tests.log gpurun_out/r2c12_timing.log
