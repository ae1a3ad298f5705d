#!/bin/bash
# round-2 GPU call 6: register-budget variants of the env kernel; launch list + ncu captures of the round-2 kernels
mkdir -p gpurun_out
{
for v in default w40 w48; do
  if [ $v = default ]; then unset DDQC_LIB; else export DDQC_LIB=tools/variants/lib_$v.so; fi
  echo "== $v"; python tools/env_launch_gap.py 2>&1 | tail -1
done
} > gpurun_out/r2c6_env_variants.log 2>&1
unset DDQC_LIB
T=r2b
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --ddmpc-steps 3 --no-graph --no-grid --no-small --no-strong > gpurun_out/ncu_l_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step -s 10 -c 1 -o gpurun_out/prof_env_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-policy --no-e2e --no-grid --no-small --no-strong > gpurun_out/ncu_e_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_finalize -s 10 -c 1 -o gpurun_out/prof_envfin_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-policy --no-e2e --no-grid --no-small --no-strong > gpurun_out/ncu_f_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_mlp -s 3 -c 1 -o gpurun_out/prof_policy_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-e2e --no-grid --no-small --no-strong > gpurun_out/ncu_p_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ddmpc_solve -s 3 -c 1 -o gpurun_out/prof_mpc_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-policy --no-e2e --ddmpc-steps 4 --no-grid --no-small --no-strong > gpurun_out/ncu_m_$T.log 2>&1
cat gpurun_out/r2c6_env_variants.log; tail -2 gpurun_out/ncu_l_$T.log gpurun_out/ncu_m_$T.log | cut -c1-300
