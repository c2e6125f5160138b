#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for n in 2 3; do
echo "== quick bench C2 cta$n"; ACMMP_B200_LIB=$PWD/acmmp-spherical_b200/lib/libacmmp_b200_cta$n.so timeout 900 python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --no-ref --out gpurun_out/quick_c2_cta$n.json > gpurun_out/quick_c2_cta$n.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/quick_c2_cta$n.log
done
