#!/bin/bash
# round-2 GPU call 29 (r2j): bench line after the last policy-kernel change, ncu capture of the policy kernel, policy + comparison tests
mkdir -p gpurun_out
T=r2j
python bench.py --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
ncu --set full --clock-control none --import-source on -k regex:policy_mlp -s 3 -c 1 -o gpurun_out/prof_policy_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-e2e --no-grid --no-small --no-strong > gpurun_out/ncu_p_$T.log 2>&1
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_comparison_gpu.py tests/test_grid_search_gpu.py -q -m gpu 2>&1 | tail -3
tail -c 200 gpurun_out/bench_$T.err; head -c 300 gpurun_out/bench_$T.json; echo
