#!/bin/bash
# records after the device pool / Park change: full GPU suite, bench C2 (with the C++ driver leg) and C4
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2n_pytest_gpu.log; tail -5 gpurun_out/r2n_pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2n_bench_c2.json 2> gpurun_out/r2n_bench_c2.err; echo "c2 rc=$?"
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r2n_bench_c4.json 2> gpurun_out/r2n_bench_c4.err; echo "c4 rc=$?"
for f in c2 c4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2n_bench_$f.json").read().replace("NaN","null"))
    print("$f", {k:d.get(k) for k in ("value","ms_per_step")}, d.get("e2e",{}).get("value"), d.get("ms_per_checkerboard_pass"), (d.get("roofline") or {}).get("frac"), d.get("e2e_driver"), d.get("gpu_launches"))
except Exception as e: print("$f", "unreadable", e)
PY
done
nvidia-smi -L
