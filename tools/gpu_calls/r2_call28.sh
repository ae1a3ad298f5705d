#!/bin/bash
# policy MLP: deferred action store A/B + policy parity tests
mkdir -p gpurun_out
{
for v in default noahead default noahead; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  python tools/policy_variants.py 2>&1 | tail -1
done
} > gpurun_out/r2_call28_policy_variants.log 2>&1
unset DDQC_LIB
cat gpurun_out/r2_call28_policy_variants.log
timeout 600 python -m pytest tests/test_policy_gpu.py tests/test_comparison_gpu.py -q -m gpu 2>&1 | tail -5
