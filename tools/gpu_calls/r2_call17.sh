#!/bin/bash
# round-2 GPU call 17: policy MLP - four tanh per reciprocal (5 MUFU per 4 tanh instead of 8): accuracy + A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_policy_gpu.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2c17_policy.log
{
for v in default tanh2 default tanh2; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== $v"; python tools/policy_variants.py 2>&1 | tail -1
done
} >> gpurun_out/r2c17_policy.log 2>&1
cat gpurun_out/r2c17_policy.log
