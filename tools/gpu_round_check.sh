python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests_s4a.log
python bench.py > gpurun_out/bench_s4a.json 2> gpurun_out/bench_s4a.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_s4a.json 2> gpurun_out/bench_ref_s4a.err
for v in default pf1 pf2 pf3 default pf1 pf2 pf3; do if [ $v = default ]; then python tools/env_perf.py; else DDQC_LIB=tools/variants/lib_$v.so python tools/env_perf.py; fi; done > gpurun_out/env_variants_s4a.log 2>&1
python tools/mpc_perf.py 592 > gpurun_out/mpc_perf_s4a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_s4a.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --ddmpc-steps 3 --no-graph > gpurun_out/ncu_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step -s 10 -c 1 -o gpurun_out/prof_env_s4a -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-policy --no-e2e > gpurun_out/ncu_e4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_mlp -s 3 -c 1 -o gpurun_out/prof_policy_s4a -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ddmpc --no-e2e > gpurun_out/ncu_p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ddmpc_solve -s 2 -c 1 -o gpurun_out/prof_mpc_s4a -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-policy --no-e2e --ddmpc-steps 3 > gpurun_out/ncu_m4.log 2>&1
tail -3 gpurun_out/gpu_tests_s4a.log; cat gpurun_out/env_variants_s4a.log; tail -2 gpurun_out/mpc_perf_s4a.log
