#!/bin/bash
# pose-first physics + restore fix: full GPU suite (no -x), env kernel timing
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_call22_tests.log
cat gpurun_out/r2_call22_tests.log
