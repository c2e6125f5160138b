#!/bin/bash
python -m pytest tests -m gpu -q -k "sphere or pruning or prior_and_hierarchy or fusion" 2>&1 | tail -25 > gpurun_out/r2e_pytest_gpu.log; tail -8 gpurun_out/r2e_pytest_gpu.log
for tp in 0 5.96e-8; do
  python tools/quick_bench.py --model sphere --width 3200 --height 1600 --views 9 --no-ref --tap-prune $tp --out gpurun_out/quick_c4_r2e_prune_$tp.json > gpurun_out/quick_c4_r2e_prune_$tp.log 2>&1
  echo "prune $tp rc=$?"; tail -1 gpurun_out/quick_c4_r2e_prune_$tp.log | cut -c1-420
done
tools/gpu_ncu_pass.sh r2e_sphere default sphere
