#!/bin/bash
mkdir -p gpurun_out
echo "== bench b200"; timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench_b200.json; tail -5 gpurun_out/bench_b200.err
echo "== bench reference"; timeout 1500 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_ref.json | grep -v '^iteration' ; tail -5 gpurun_out/bench_ref.err
