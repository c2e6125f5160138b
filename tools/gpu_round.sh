#!/bin/bash
# Round-end GPU session: tests, smoke, both bench arms, ncu launch list of the bench command.
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench b200"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b200.json').read())
print({k:d[k] for k in ('value','ms_per_step','e2e','ms_per_checkerboard_pass','gpu_launches','depth_within_1pct_of_ground_truth','clocks')}); print(d.get('roofline',{}).get('frac'), d.get('cpu_baseline'))
PY
tail -3 gpurun_out/bench_b200.err
echo "== bench reference"; timeout 1500 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','ms_per_checkerboard_pass','depth_within_1pct_of_ground_truth')})
PY
echo "== ncu launch list of the bench command"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/launches_bench.csv
