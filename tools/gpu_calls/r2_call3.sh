#!/bin/bash
# round-2 GPU call 3: early-load env kernel (parity + timing), new bench.py end to end
mkdir -p gpurun_out
python -m pytest tests/test_env_parity_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2c3_env_tests.log
{
python tools/env_launch_gap.py 2>&1 | tail -1
python tools/env_launch_gap.py 131072 2>&1 | tail -1
} > gpurun_out/r2c3_env_timing.log 2>&1
python bench.py > gpurun_out/bench_r2c3.json 2> gpurun_out/bench_r2c3.err
python bench.py --impl reference > gpurun_out/bench_ref_r2c3.json 2> gpurun_out/bench_ref_r2c3.err
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c3_gpu_tests.log
cat gpurun_out/r2c3_env_tests.log gpurun_out/r2c3_env_timing.log gpurun_out/r2c3_gpu_tests.log; tail -c 1500 gpurun_out/bench_r2c3.err; head -c 600 gpurun_out/bench_r2c3.json
