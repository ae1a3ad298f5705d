#!/bin/bash
# dual-recovery kernel, left-looking factorisation: parity of the optional modes + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddmpc_parity_gpu.py tests/test_ddmpc_eval_gpu.py -q -m gpu -k "previous or threshold or optional" 2>&1 | tail -4
python tools/recover_perf.py 2>&1 | tail -1 | tee gpurun_out/r2_call32_recover_perf.log
python tools/recover_perf.py 2>&1 | tail -1 | tee -a gpurun_out/r2_call32_recover_perf.log
