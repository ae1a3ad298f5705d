#!/bin/bash
# round-2 GPU call 4 (2 GPUs): multi-rank bench (strong-scaling leg, sharded grid search)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2c4_2gpu.json 2> gpurun_out/bench_r2c4_2gpu.err
tail -c 2000 gpurun_out/bench_r2c4_2gpu.err; head -c 400 gpurun_out/bench_r2c4_2gpu.json
python -m pytest tests/test_env_parity_gpu.py -q -m gpu -k non_current 2>&1 | tail -3
