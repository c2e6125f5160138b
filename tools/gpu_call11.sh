#!/bin/bash
mkdir -p gpurun_out
QB="python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --no-ref"
timeout 600 $QB > gpurun_out/quick_c2b.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_pass' -s 4 -c 2 -o gpurun_out/prof_pass_r1f $QB > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
