#!/bin/bash
python -m pytest tests -m gpu -q -k "sphere or pruning or twin or equirect" 2>&1 | tail -25 > gpurun_out/r2g_pytest_gpu.log; tail -6 gpurun_out/r2g_pytest_gpu.log
python tools/quick_bench.py --model sphere --width 3200 --height 1600 --views 9 --no-ref --out gpurun_out/quick_c4_r2g_ring.json > gpurun_out/quick_c4_r2g_ring.log 2>&1; tail -1 gpurun_out/quick_c4_r2g_ring.log | cut -c1-400
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/r2g_bench_c4.err | cut -c1-300
python bench.py --config C3 --steps 1 --warmup 3 > gpurun_out/r2g_bench_c3_n1.json 2> gpurun_out/r2g_bench_c3_n1.err; echo "c3 rc=$?"; tail -3 gpurun_out/r2g_bench_c3_n1.err | cut -c1-300
for f in c4 c3_n1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_$f.json").read().replace("NaN","null"))
    print("$f", {k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling")}, d.get("e2e",{}).get("value"), d.get("ms_per_checkerboard_pass"), d.get("sphere_tap_pruning"), d.get("depth_within_1pct_of_ground_truth"), d.get("prior_host_s_not_hidden_per_step"), (d.get("roofline") or {}).get("frac"))
except Exception as e: print("$f", "unreadable", e)
PY
done
