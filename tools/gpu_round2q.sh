#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2q_pytest_gpu.log; tail -5 gpurun_out/r2q_pytest_gpu.log
ACMMP_TRACE=1 python tools/driver_bench.py --views 11 --skip-files --no-fusion --trace gpurun_out/r2q_trace --out gpurun_out/r2q_driver.json > gpurun_out/r2q_driver.log 2>&1; echo "driver rc=$?"; tail -1 gpurun_out/r2q_driver.log | cut -c1-1800
grep -h "pool\.\|reserve\|prior_from_triangles \|set_views \|next_level " gpurun_out/r2q_trace_resident_gpu_prior.txt
ACMMP_TRACE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_c2_trace.json 2> gpurun_out/r2q_bench_c2_trace.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2q_bench_c2_trace.json").read().replace("NaN","null"))
e=d.get("e2e_driver") or {}
print(e.get("s_per_view"), e.get("breakdown_s"))
print("\n".join(l for l in e.get("trace", []) if "pool." in l or "reserve" in l or "prior_from_triangles " in l or "set_views " in l or "next_level " in l))
PY
