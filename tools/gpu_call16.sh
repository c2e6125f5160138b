#!/bin/bash
mkdir -p gpurun_out
echo "== quick bench C4 sphere"; timeout 900 python tools/quick_bench.py --model sphere --width 3200 --height 1600 --views 9 --out gpurun_out/quick_c4.json > gpurun_out/quick_c4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/quick_c4.log | cut -c1-1500
