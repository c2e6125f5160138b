#!/bin/bash
# A/B of library variants on the geometric stage (and the photometric one) at C2 / C4 size: tools/gpu_ab_geom.sh <tag> <variant> ...
tag=$1; shift
mkdir -p gpurun_out
for model in pinhole sphere; do
if [ "$model" = sphere ]; then args="--model sphere --width 3200 --height 1600 --views 9 --no-ref --geom"
else args="--width 3200 --height 2130 --focal 2800 --views 11 --no-ref --geom"; fi
for v in "$@"; do
  lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200_$v.so
  [ "$v" = default ] && lib=$PWD/acmmp-spherical_b200/lib/libacmmp_b200.so
  ACMMP_B200_LIB=$lib timeout 600 python tools/quick_bench.py $args --out gpurun_out/quick_${model}_${tag}_$v.json > gpurun_out/quick_${model}_${tag}_$v.log 2>&1
  python - <<PY
import json
d=json.load(open("gpurun_out/quick_${model}_${tag}_$v.json"))
print("$model $v photometric %.2f geom %.2f / %.2f ms per pass, checksum %r" % (d["mine_run1"]["pass_sum_ms"]/6, d["geom_run0"]["pass_sum_ms"]/d["geom_run0"]["n_pass"], d["geom_run1"]["pass_sum_ms"]/d["geom_run1"]["n_pass"], d["geom_checksum"]))
PY
done
done
