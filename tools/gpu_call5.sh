#!/bin/bash
mkdir -p gpurun_out
tools/texbench/tex_bench > gpurun_out/tex_bench.jsonl 2>&1; cat gpurun_out/tex_bench.jsonl
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== quick bench half"; timeout 600 python tools/quick_bench.py --width 1600 --height 1065 --focal 1400 --views 11 --out gpurun_out/quick_half.json > gpurun_out/quick_half.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/quick_half.log
echo "== quick bench C2"; timeout 900 python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --out gpurun_out/quick_c2.json > gpurun_out/quick_c2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/quick_c2.log
