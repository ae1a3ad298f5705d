#!/bin/bash
# round-2 GPU call 13: state of the tree after the env-kernel experiments: full GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c13_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c13_smoke.log 2>&1
python bench.py > gpurun_out/bench_r2c13.json 2> gpurun_out/bench_r2c13.err
cat gpurun_out/r2c13_gpu_tests.log; tail -3 gpurun_out/r2c13_smoke.log; tail -c 400 gpurun_out/bench_r2c13.err; head -c 300 gpurun_out/bench_r2c13.json
