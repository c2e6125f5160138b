#!/bin/bash
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r2ac_bench_c4.json 2> gpurun_out/r2ac_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ac_bench_c4.json").read().replace("NaN","null"))
print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["roofline"]["frac"], d.get("e2e_driver"))
PY
