#!/bin/bash
# `ncu --set full` of the small kernels of a stage at C2 size (k_random_init in the quad form, median filter, depth / normal
# conversion), after the same command ran clean without ncu
mkdir -p gpurun_out
args="--width 3200 --height 2130 --focal 2800 --views 11 --no-ref"
timeout 600 python tools/quick_bench.py $args > gpurun_out/plain_small.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_random_init|k_median_filter|k_depth_normal|k_rng_fill|k_pad_reference" -c 6 -f \
    -o gpurun_out/prof_small_r2 python tools/quick_bench.py $args > gpurun_out/ncu_small.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_small.log | cut -c1-300
