#!/bin/bash
# round-2 GPU call 2: env cache-hint variants, fused-vs-split finalisation at small batches (graph replay), ncu of the split kernel
mkdir -p gpurun_out
python -m pytest tests/test_env_parity_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2c2_env_tests.log
{
for v in default st1 st2 st2ld1 ld1 t512; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== $v"
  python tools/env_launch_gap.py 2>&1 | tail -1
done
unset DDQC_LIB
for n in 4096 32768 131072 262144; do
  for f in 0 100000000; do
    DDQC_ENV_FUSED_MAX_N=$f python tools/env_launch_gap.py $n 2>&1 | tail -1
  done
done
} > gpurun_out/r2c2_env_variants.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step -s 10 -c 1 -o gpurun_out/prof_env_r2a -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-policy --no-e2e > gpurun_out/ncu_e_r2a.log 2>&1
cat gpurun_out/r2c2_env_tests.log gpurun_out/r2c2_env_variants.log; tail -3 gpurun_out/ncu_e_r2a.log
