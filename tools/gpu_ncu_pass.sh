#!/bin/bash
# One `ncu --set full` capture of a late k_pass launch of the quick bench, after the same command ran clean without ncu:
#   tools/gpu_ncu_pass.sh <tag> [variant] [pinhole|sphere]
tag=$1; v=${2:-default}; model=${3:-pinhole}
lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200_$v.so
[ "$v" = default ] && lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200.so
mkdir -p gpurun_out
if [ "$model" = sphere ]; then args="--model sphere --width 3200 --height 1600 --views 9 --no-ref"
else args="--width 3200 --height 2130 --focal 2800 --views 11 --no-ref"; fi
ACMMP_B200_LIB=$lib timeout 600 python tools/quick_bench.py $args > gpurun_out/plain_${tag}_$v.log 2>&1 &&
ACMMP_B200_LIB=$lib timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_pass -s 4 -c 1 -f \
    -o gpurun_out/prof_pass_${tag}_$v python tools/quick_bench.py $args > gpurun_out/ncu_${tag}_$v.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_${tag}_$v.log | cut -c1-300
