#!/bin/bash
# round-2 GPU call 10: pair kernel (two envs per thread, packed fp32x2): parity + timing + ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_env_tile_kernel_gpu.py -x -q -m gpu -k "pair or default_threshold" 2>&1 | tail -15 > gpurun_out/r2c10_pair_tests.log
timeout 600 python -m pytest tests/test_env_parity_gpu.py -x -q -m gpu -k "baseline_sizes" 2>&1 | tail -5 >> gpurun_out/r2c10_pair_tests.log
{
echo "== pair kernel (default)"; timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
echo "== plain kernel"; DDQC_ENV_PAIR_MIN_N=100000000000 timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
for n in 131072 262144; do
  echo "== N=$n pair"; timeout 120 python tools/env_launch_gap.py $n 2>&1 | tail -1
  echo "== N=$n plain"; DDQC_ENV_PAIR_MIN_N=100000000000 timeout 120 python tools/env_launch_gap.py $n 2>&1 | tail -1
done
} > gpurun_out/r2c10_pair_timing.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_pair -s 10 -c 1 -o gpurun_out/prof_envpair_r2e -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-policy --no-e2e --no-grid --no-small --no-strong > gpurun_out/ncu_pair_r2e.log 2>&1
cat gpurun_out/r2c10_pair_tests.log gpurun_out/r2c10_pair_timing.log
