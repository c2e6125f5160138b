#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== bench N=8"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b200_n8.json 2> gpurun_out/bench_b200_n8.err; echo "rc=$?"; tail -c 1800 gpurun_out/bench_b200_n8.json; tail -5 gpurun_out/bench_b200_n8.err
