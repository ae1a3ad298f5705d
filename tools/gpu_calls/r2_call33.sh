#!/bin/bash
# ncu capture of the left-looking dual-recovery kernel
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ddmpc_recover -s 2 -c 1 -o gpurun_out/prof_rec_r2k -f python tools/recover_perf.py 1184 > gpurun_out/ncu_rec_r2k.log 2>&1
tail -2 gpurun_out/ncu_rec_r2k.log | cut -c1-200
