#!/bin/bash
# developer build of libddqc.so with extra -D flags for the DD-MPC unit (e.g. -DDDQC_PROF_FACTOR)
set -e
D=/root/repo/data_driven_quad_control_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I /root/repo/include "$@" -c $D/ddqc_ddmpc.cu -o $D/ddqc_ddmpc.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/data_driven_quad_control_b200/libddqc.so $D/ddqc_env.o $D/ddqc_env_aux.o $D/ddqc_ctrl.o $D/ddqc_ddmpc.o $D/ddqc_ddmpc_recover.o $D/ddqc_policy.o -cudart static
