#!/bin/bash
# developer probe: builds libddqc.so variants of the dual-recovery unit, e.g. tools/variants/build_rec.sh b3 -DDDQC_REC_BLOCKS=3
set -e
name=$1; shift
R=/root/repo; D=$R/data_driven_quad_control_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I $R/include "$@" -c $D/ddqc_ddmpc_recover.cu -o /tmp/ddqc_rec_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $R/tools/variants/lib_$name.so $D/ddqc_env.o $D/ddqc_env_aux.o $D/ddqc_ctrl.o $D/ddqc_ddmpc.o /tmp/ddqc_rec_$name.o $D/ddqc_policy.o -cudart static
echo built tools/variants/lib_$name.so
