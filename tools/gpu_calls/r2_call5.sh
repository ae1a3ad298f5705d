#!/bin/bash
# round-2 GPU call 5: full GPU test suite (ground watch, alpha_sr anchor mode), FP64 peak micro-benchmark with its clock record
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2c5_gpu_tests.log
{
echo "# tools/ubench_fp64 (builder's FP64 peak micro-benchmark; the DD-MPC roofline denominator) on $(nvidia-smi --query-gpu=name --format=csv,noheader)"
echo "# clocks before: $(nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv,noheader)"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_fp64 tools/ubench_fp64.cu 2>&1 | tail -2
/tmp/ubench_fp64 & pid=$!
sleep 0.5; echo "# clocks under load: $(nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv,noheader)"
wait $pid
echo "# clocks after: $(nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv,noheader)"
} > gpurun_out/r2_ubench_fp64.txt 2>&1
cat gpurun_out/r2c5_gpu_tests.log gpurun_out/r2_ubench_fp64.txt
