#!/bin/bash
python -m pytest tests/test_gpu_cpp_driver.py tests/test_gpu_edge_cases.py -m gpu -q -k "cpp_driver or reserved or parked" 2>&1 | tail -6 > gpurun_out/try3.log; tail -3 gpurun_out/try3.log | cut -c1-600
python tools/driver_bench.py --views 11 --skip-files --out gpurun_out/r2ab_driver.json > gpurun_out/r2ab_driver.log 2>&1; echo "driver rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ab_driver.json"))
e=d["resident_gpu_prior"]; print(e["s_per_view"], {x:e.get(x) for x in ("wall_s","setup_s","views_s","run_s","output_s","sweep1_s","geom_s","fusion_s","process_wall_s")})
PY
