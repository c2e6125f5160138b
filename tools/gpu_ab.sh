#!/bin/bash
# A/B timing of library variants on one box: tools/gpu_ab.sh <tag> <variant> [<variant> ...]
# variant "" or "default" = lib/libacmmp_b200.so, otherwise lib/libacmmp_b200_<variant>.so (built with extra -D flags).
# Variants must run in the SAME call: boxes differ by a few percent.
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200_$v.so
  [ "$v" = default ] && lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200.so
  echo "== quick bench C2 $v"
  ACMMP_B200_LIB=$lib timeout 600 python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --no-ref \
      --out gpurun_out/quick_c2_${tag}_$v.json > gpurun_out/quick_c2_${tag}_$v.log 2>&1
  echo "rc=$?"; tail -1 gpurun_out/quick_c2_${tag}_$v.log | cut -c150-300
done
