#!/bin/bash
# A/B timing of library variants on the C4-size equirectangular stage (3200x1600, 8 source views), same box:
#   tools/gpu_ab_sphere.sh <tag> <variant> [<variant> ...]      (variants as in tools/gpu_ab.sh)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200_$v.so
  [ "$v" = default ] && lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200.so
  echo "== quick bench C4 $v"
  ACMMP_B200_LIB=$lib timeout 900 python tools/quick_bench.py --model sphere --width 3200 --height 1600 --views 9 --no-ref \
      --out gpurun_out/quick_c4_${tag}_$v.json > gpurun_out/quick_c4_${tag}_$v.log 2>&1
  echo "rc=$?"; tail -1 gpurun_out/quick_c4_${tag}_$v.log | cut -c1-400
done
