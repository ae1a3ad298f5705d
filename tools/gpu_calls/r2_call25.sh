#!/bin/bash
# A/B: env substep loop (deferred velocity half vs sequential), DD-MPC solver with the phase routines out of line
mkdir -p gpurun_out
{
for v in default seqv default seqv; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== $v"; python tools/env_launch_gap.py 2>&1 | tail -1
done
} > gpurun_out/r2_call25_env_variants.log 2>&1
unset DDQC_LIB
cat gpurun_out/r2_call25_env_variants.log | cut -c1-300
tools/variants/run.sh base outline base outline > gpurun_out/r2_call25_mpc_variants.log 2>&1
cat gpurun_out/r2_call25_mpc_variants.log
