#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench b200"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "rc=$?"; cat gpurun_out/bench_b200.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','ms_per_checkerboard_pass','gpu_launches','depth_within_1pct_of_ground_truth','clocks')}); print(d.get('roofline',{}).get('frac'), d.get('cpu_baseline'))"; tail -3 gpurun_out/bench_b200.err
