#!/bin/bash
# final check of the round: full GPU suite, smoke, default bench
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2y_pytest_gpu.log; tail -3 gpurun_out/r2y_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/r2y_bench_c2.json 2> gpurun_out/r2y_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2y_bench_c2.json").read().replace("NaN","null"))
e=d.get("e2e_driver") or {}
print({k:d.get(k) for k in ("value","ms_per_step","steps","warmup","gpu_launches")}, "e2e", d["e2e"]["value"], d["roofline"]["frac"], e.get("s_per_view"), e.get("fusion_s"), e.get("breakdown_s"))
PY
