#!/bin/bash
# round-2 GPU call 27 (r2i): full GPU suite, smoke(), bench lines of both arms, launch list, ncu capture of the policy kernel
mkdir -p gpurun_out
T=r2i
timeout 2700 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/tests_$T.log; cat gpurun_out/tests_$T.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/smoke_$T.log; cat gpurun_out/smoke_$T.log
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python bench.py --impl reference > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --ddmpc-steps 3 --no-graph --no-grid --no-small --no-strong --no-e2e > gpurun_out/ncu_l_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_mlp -s 3 -c 1 -o gpurun_out/prof_policy_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-e2e --no-grid --no-small --no-strong > gpurun_out/ncu_p_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step -s 10 -c 1 -o gpurun_out/prof_env_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-policy --no-e2e --no-grid --no-small --no-strong > gpurun_out/ncu_e_$T.log 2>&1
tail -c 300 gpurun_out/bench_$T.err; head -c 500 gpurun_out/bench_$T.json; echo
