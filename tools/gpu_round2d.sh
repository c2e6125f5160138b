#!/bin/bash
# round-2 GPU call d: tests, sphere A/B (ping-pong), bench C2 (both arms) + C4 (both arms), C3 smoke at N=1 with a small step count
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2d_pytest_gpu.log; tail -6 gpurun_out/r2d_pytest_gpu.log
tools/gpu_ab_sphere.sh r2d r2c default 2>&1 | cut -c1-330
python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err; echo "c2 rc=$?"; tail -3 gpurun_out/r2d_bench_c2.err
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r2d_bench_c2_ref.json 2> gpurun_out/r2d_bench_c2_ref.err; echo "c2 ref rc=$?"
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err; echo "c4 rc=$?"; tail -3 gpurun_out/r2d_bench_c4.err
python bench.py --config C4 --impl reference --steps 2 --warmup 3 > gpurun_out/r2d_bench_c4_ref.json 2> gpurun_out/r2d_bench_c4_ref.err; echo "c4 ref rc=$?"; tail -3 gpurun_out/r2d_bench_c4_ref.err
for f in c2 c2_ref c4 c4_ref; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2d_bench_$f.json").read().replace("NaN","null"))
    print("$f", {k:d.get(k) for k in ("value","ms_per_step")}, d.get("e2e",{}).get("value"), d.get("ms_per_checkerboard_pass"), (d.get("roofline") or {}).get("frac"), d.get("e2e_driver"))
except Exception as e: print("$f", "unreadable", e)
PY
done
