#!/bin/bash
# developer probe: builds libddqc.so variants of the env-step unit, e.g.
#   tools/variants/build_env.sh fin0 -DDDQC_ENV_FINALIZE=0
# -> tools/variants/lib_fin0.so (the other units are taken from the in-tree build)
set -e
name=$1; shift
R=/root/repo; D=$R/data_driven_quad_control_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I $R/include -fmad=false "$@" -c $D/ddqc_env.cu -o /tmp/ddqc_env_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $R/tools/variants/lib_$name.so /tmp/ddqc_env_$name.o $D/ddqc_env_aux.o $D/ddqc_ctrl.o $D/ddqc_ddmpc.o $D/ddqc_ddmpc_recover.o $D/ddqc_policy.o -cudart static
echo built tools/variants/lib_$name.so
