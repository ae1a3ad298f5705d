#!/bin/bash
# round-2 GPU call 7: bulk-copy tile kernel - parity (under a timeout: a wrong mbarrier protocol hangs) and timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_tile_kernel_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2c7_tile_tests.log
echo "tile tests rc=$?" >> gpurun_out/r2c7_tile_tests.log
{
echo "== tile kernel (default)"; timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
echo "== plain kernel (DDQC_ENV_TILE_MIN_N huge)"; DDQC_ENV_TILE_MIN_N=100000000000 timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
for n in 131072 262144 524288; do
  echo "== N=$n tile"; DDQC_ENV_TILE_MIN_N=1024 timeout 120 python tools/env_launch_gap.py $n 2>&1 | tail -1
  echo "== N=$n plain"; DDQC_ENV_TILE_MIN_N=100000000000 timeout 120 python tools/env_launch_gap.py $n 2>&1 | tail -1
done
} > gpurun_out/r2c7_tile_timing.log 2>&1
timeout 600 python -m pytest tests/test_env_parity_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2c7_env_tests.log
cat gpurun_out/r2c7_tile_tests.log gpurun_out/r2c7_tile_timing.log gpurun_out/r2c7_env_tests.log
