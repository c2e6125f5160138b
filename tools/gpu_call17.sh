#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python -c "
import json
d=json.load(open('gpurun_out/metrics_ncc_sphere.json')); print({k:(round(v['frac_1e4'],4),round(v['frac_1e3'],5)) for k,v in d.items()})
d=json.load(open('gpurun_out/metrics_full_stage_sphere.json')); print(d)"
echo "== quick bench C4 sphere"; timeout 900 python tools/quick_bench.py --model sphere --width 3200 --height 1600 --views 9 --no-ref --out gpurun_out/quick_c4.json > gpurun_out/quick_c4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/quick_c4.log | cut -c1-600
