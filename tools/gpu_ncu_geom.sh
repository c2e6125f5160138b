#!/bin/bash
# One `ncu --set full` capture of a full-resolution GEOMETRIC k_pass launch of the bench (the 21st k_pass<.,geom> launch of
# the first step is the 5th geometric pass of the finest level): tools/gpu_ncu_geom.sh <tag>
tag=$1
mkdir -p gpurun_out
timeout 1200 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:k_passILi0ELi2 -s 20 -c 1 -f \
    -o gpurun_out/prof_pass_geom_$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_geom_$tag.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_geom_$tag.log | cut -c1-200
