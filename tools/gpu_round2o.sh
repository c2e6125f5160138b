#!/bin/bash
tools/gpu_ab_env.sh r2o none ACMMP_TILE_ORDER=1
ACMMP_TRACE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_c2_trace.json 2> gpurun_out/r2o_bench_c2_trace.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2o_bench_c2_trace.json").read().replace("NaN","null"))
e=d.get("e2e_driver") or {}
print(e.get("s_per_view"), e.get("breakdown_s"))
print("\n".join(e.get("trace", [])))
PY
