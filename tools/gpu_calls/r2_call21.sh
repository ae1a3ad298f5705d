#!/bin/bash
# pose-first physics: full GPU suite, the comparison run against the digitised reference figure, env step timing
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2_call21_tests.log
cat gpurun_out/r2_call21_tests.log
timeout 900 python tools/comparison_vs_plot.py > gpurun_out/r2_call21_plot.log 2>&1; tail -8 gpurun_out/r2_call21_plot.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-ddmpc --no-policy --no-cpu-baseline --no-e2e --no-small --no-strong --no-grid > gpurun_out/r2_call21_bench.json 2> gpurun_out/r2_call21_bench.err; tail -c 1500 gpurun_out/r2_call21_bench.json
