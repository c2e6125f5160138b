#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3 4 5 6 7 8; do timeout 60 tools/tma/tma_test $v >> gpurun_out/tma_variants.log 2>&1; echo "variant $v rc=$?" >> gpurun_out/tma_variants.log; done
cat gpurun_out/tma_variants.log
export ACMMP_NO_TMA=1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== quick bench C1"; timeout 600 python tools/quick_bench.py --width 640 --height 480 --focal 500 --views 5 --out gpurun_out/quick_c1.json > gpurun_out/quick_c1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/quick_c1.log
echo "== quick bench C2"; timeout 900 python tools/quick_bench.py --width 3200 --height 2130 --focal 2800 --views 11 --out gpurun_out/quick_c2.json > gpurun_out/quick_c2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/quick_c2.log
