#!/bin/bash
# One `ncu --set full` capture of a late k_pass launch of the quick bench: tools/gpu_ncu_pass.sh <tag> [variant]
tag=$1; v=${2:-default}
lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200_$v.so
[ "$v" = default ] && lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200.so
mkdir -p gpurun_out
ACMMP_B200_LIB=$lib timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_pass -s 4 -c 1 -f \
    -o gpurun_out/prof_pass_${tag}_$v python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --no-ref \
    > gpurun_out/ncu_${tag}_$v.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_${tag}_$v.log | cut -c1-300
