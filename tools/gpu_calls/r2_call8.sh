#!/bin/bash
# round-2 GPU call 8: tile kernel variants (tile size) and an ncu capture of the tile kernel
mkdir -p gpurun_out
{
for v in tile128 tile512; do
  export DDQC_LIB=tools/variants/lib_$v.so
  echo "== $v"; timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
done
} > gpurun_out/r2c8_tile_variants.log 2>&1
unset DDQC_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_tile -s 10 -c 1 -o gpurun_out/prof_envtile_r2c -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-policy --no-e2e --no-grid --no-small --no-strong > gpurun_out/ncu_t_r2c.log 2>&1
cat gpurun_out/r2c8_tile_variants.log; tail -3 gpurun_out/ncu_t_r2c.log | cut -c1-300
