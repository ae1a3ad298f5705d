#!/bin/bash
# round-2 GPU call 15: final 1-GPU bench lines (both arms), launch list of the whole bench
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
python bench.py --impl reference > gpurun_out/bench_ref_r2g.json 2> gpurun_out/bench_ref_r2g.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --ddmpc-steps 3 --no-graph --no-grid --no-small --no-strong --no-e2e > gpurun_out/ncu_l_r2g.log 2>&1
tail -c 300 gpurun_out/bench_r2g.err; head -c 400 gpurun_out/bench_r2g.json; echo; head -c 600 gpurun_out/bench_ref_r2g.json; echo; tail -2 gpurun_out/ncu_l_r2g.log | cut -c1-200
