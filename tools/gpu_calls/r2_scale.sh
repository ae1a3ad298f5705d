#!/bin/bash
# multi-GPU bench line: tools/r2_scale.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/bench_r2i_${N}gpu.out 2> gpurun_out/bench_r2i_${N}gpu.err
grep '^{' gpurun_out/bench_r2i_${N}gpu.out > gpurun_out/bench_r2i_${N}gpu.json
tail -c 600 gpurun_out/bench_r2i_${N}gpu.err; head -c 300 gpurun_out/bench_r2i_${N}gpu.json
