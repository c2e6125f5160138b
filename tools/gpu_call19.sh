#!/bin/bash
mkdir -p gpurun_out
QB="python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --no-ref"
timeout 900 ncu --set full --clock-control none -k regex:'k_depth_normal|k_median_filter|k_random_init|k_pad_reference' -c 5 -o gpurun_out/prof_small_r1i $QB > gpurun_out/ncu_small.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_small.log
