#!/bin/bash
# parity tests on a library variant, then A/B timing: tools/gpu_try.sh <tag> <variant-under-test> [baseline-variant]
tag=$1; v=$2; base=${3:-default}
mkdir -p gpurun_out
echo "== pytest gpu on $v"
ACMMP_B200_LIB=$PWD/acmmp-spherical_b200/lib/libacmmp_b200_$v.so timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu_$v.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$v.log
bash tools/gpu_ab.sh $tag $base $v
