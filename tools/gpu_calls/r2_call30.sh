#!/bin/bash
# policy kernel with programmatic dependent launch: rollout timing A/B (eager and inside bench's graph), policy + env tests
mkdir -p gpurun_out
{
for v in default nopdl default nopdl; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== $v (default = PDL wait only; nopdl = plain launch)"
  python tools/policy_perf.py 2>&1 | tail -1
  python bench.py --steps 100 --warmup 10 --no-ddmpc --no-cpu-baseline --no-e2e --no-small --no-strong --no-grid 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['policy_rollout']; print('bench graph rollout ms/step', p['ms_per_step'], 'policy fwd ms', p['policy_forward_ms'], 'env ms', d['ms_per_step'])"
done
} > gpurun_out/r2_call30_policy_pdl.log 2>&1
unset DDQC_LIB
cat gpurun_out/r2_call30_policy_pdl.log | cut -c1-400
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_comparison_gpu.py -q -m gpu 2>&1 | tail -3
