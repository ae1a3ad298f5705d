#!/bin/bash
# round-2 GPU call 18: clipf as NaN-propagating min/max (A/B against the select form) + env parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_env_parity_gpu.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2c18_clip.log
{
for v in default clipsel default clipsel; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== $v"; timeout 120 python tools/env_launch_gap.py 2>&1 | tail -1
done
} >> gpurun_out/r2c18_clip.log 2>&1
cat gpurun_out/r2c18_clip.log
