#!/bin/bash
# developer probe: builds libddqc.so variants of the policy-MLP unit, e.g.
#   tools/variants/build_policy.sh nopipe -DDDQC_POLICY_TMEM_PIPE=0
# -> tools/variants/lib_<name>.so (the other units are taken from the in-tree build)
set -e
name=$1; shift
R=/root/repo; D=$R/data_driven_quad_control_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I $R/include "$@" -c $D/ddqc_policy.cu -o /tmp/ddqc_policy_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $R/tools/variants/lib_$name.so $D/ddqc_env.o $D/ddqc_env_aux.o $D/ddqc_ctrl.o $D/ddqc_ddmpc.o $D/ddqc_ddmpc_recover.o /tmp/ddqc_policy_$name.o -cudart static
echo built tools/variants/lib_$name.so
