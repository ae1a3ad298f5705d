# full check of the round: GPU tests, smoke, both bench arms, launch list and one ncu --set full capture per kernel
T=${1:-s5a}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests_$T.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$T.log 2>&1
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --ddmpc-steps 3 --no-graph > gpurun_out/ncu_l_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step -s 10 -c 1 -o gpurun_out/prof_env_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-policy --no-e2e > gpurun_out/ncu_e_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_mlp -s 3 -c 1 -o gpurun_out/prof_policy_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-e2e > gpurun_out/ncu_p_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ddmpc_solve -s 3 -c 1 -o gpurun_out/prof_mpc_$T -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-policy --no-e2e --ddmpc-steps 4 > gpurun_out/ncu_m_$T.log 2>&1
tail -3 gpurun_out/gpu_tests_$T.log; tail -2 gpurun_out/smoke_$T.log; tail -c 600 gpurun_out/bench_$T.err
