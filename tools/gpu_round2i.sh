#!/bin/bash
# full GPU suite, bench C2 (with the C++ driver leg) and C4, launch list of the bench command, ncu captures of both passes
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2i_pytest_gpu.log; tail -8 gpurun_out/r2i_pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err; echo "c2 rc=$?"
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r2i_bench_c4.json 2> gpurun_out/r2i_bench_c4.err; echo "c4 rc=$?"
for f in c2 c4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2i_bench_$f.json").read().replace("NaN","null"))
    print("$f", {k:d.get(k) for k in ("value","ms_per_step")}, d.get("e2e",{}).get("value"), d.get("ms_per_checkerboard_pass"), (d.get("roofline") or {}).get("frac"), d.get("e2e_driver"), d.get("gpu_launches"))
except Exception as e: print("$f", "unreadable", e)
PY
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-driver-leg > gpurun_out/r2i_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2i_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-driver-leg > gpurun_out/r2i_ncu_bench.log 2>&1
echo "launch list rc=$?"
tools/gpu_ncu_pass.sh r2i_sphere default sphere
tools/gpu_ncu_pass.sh r2i_pinhole default pinhole
