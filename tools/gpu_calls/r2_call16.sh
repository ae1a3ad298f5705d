#!/bin/bash
# round-2 GPU call 16: DD-MPC factorisation stage counters (developer build with -DDDQC_PROF_FACTOR)
mkdir -p gpurun_out
{
echo "== product build"; python tools/mpc_perf.py 2>&1 | tail -2
echo "== stage counters"; DDQC_LIB=tools/variants/lib_fprof.so python tools/mpc_perf.py 2>&1 | tail -3
} > gpurun_out/r2c16_mpc_prof.log 2>&1
cat gpurun_out/r2c16_mpc_prof.log
