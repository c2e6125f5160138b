#!/bin/bash
mkdir -p gpurun_out
for v in th16_c1; do
echo "== quick bench C2 $v"; ACMMP_B200_LIB=$PWD/acmmp-spherical_b200/lib/libacmmp_b200_$v.so timeout 900 python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --no-ref --out gpurun_out/quick_c2_$v.json > gpurun_out/quick_c2_$v.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/quick_c2_$v.log | cut -c1-700
done
