#!/bin/bash
# A/B of one library under different environment settings on one box: tools/gpu_ab_env.sh <tag> "<VAR=val>" ["<VAR=val>" ...]
# ("none" = no setting).  quick bench = six photometric passes at C2 from random initialisation.
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  name=$(echo "$v" | tr '= ' '__')
  echo "== quick bench C2 [$v] rep $rep"
  if [ "$v" = none ]; then
    timeout 600 python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --no-ref --out gpurun_out/quick_c2_${tag}_${name}_$rep.json > gpurun_out/quick_c2_${tag}_${name}_$rep.log 2>&1
  else
    env $v timeout 600 python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --no-ref --out gpurun_out/quick_c2_${tag}_${name}_$rep.json > gpurun_out/quick_c2_${tag}_${name}_$rep.log 2>&1
  fi
  echo "rc=$?"; tail -1 gpurun_out/quick_c2_${tag}_${name}_$rep.log | cut -c150-330
done
done
