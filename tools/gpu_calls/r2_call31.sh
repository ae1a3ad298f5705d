#!/bin/bash
# dual-recovery kernel: 2 / 3 / 4 resident CTAs per SM
mkdir -p gpurun_out
{
for v in default b3 b4 default b3 b4; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  python tools/recover_perf.py 2>&1 | tail -1
done
} > gpurun_out/r2_call31_recover_variants.log 2>&1
cat gpurun_out/r2_call31_recover_variants.log
