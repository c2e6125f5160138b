#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== quick bench C2"; timeout 900 python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --no-ref --out gpurun_out/quick_c2.json > gpurun_out/quick_c2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/quick_c2.log
