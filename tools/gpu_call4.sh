#!/bin/bash
mkdir -p gpurun_out
echo "== smoke (TMA)" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== golden"; timeout 600 python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/golden.log
QB="python tools/quick_bench.py --width 1600 --height 1065 --focal 1400 --views 11"
echo "== quick bench half-res"; timeout 600 $QB --out gpurun_out/quick_half.json > gpurun_out/quick_half.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_quick_half.csv $QB > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 $QB > gpurun_out/quick_half2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_pass|BlackPixelUpdate' -s 1 -c 3 -o gpurun_out/prof_pass_r1a $QB > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
tail -1 gpurun_out/quick_half.log
