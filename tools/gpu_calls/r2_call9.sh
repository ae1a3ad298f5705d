#!/bin/bash
# round-2 GPU call 9: tile kernel, second iteration (single issuer, cached control values): parity + timing + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_tile_kernel_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c9_tile_tests.log
{
echo "== tile kernel"; timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
echo "== plain kernel"; DDQC_ENV_TILE_MIN_N=100000000000 timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
} > gpurun_out/r2c9_tile_timing.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_tile -s 10 -c 1 -o gpurun_out/prof_envtile_r2d -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-policy --no-e2e --no-grid --no-small --no-strong > gpurun_out/ncu_t_r2d.log 2>&1
cat gpurun_out/r2c9_tile_tests.log gpurun_out/r2c9_tile_timing.log
