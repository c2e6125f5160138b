#!/bin/bash
python -m pytest tests/test_gpu_fusion.py tests/test_gpu_cpp_driver.py -m gpu -q 2>&1 | tail -12 > gpurun_out/try3.log; tail -6 gpurun_out/try3.log | cut -c1-600
python tools/driver_bench.py --views 11 --skip-files --out gpurun_out/r2w_driver.json > gpurun_out/r2w_driver.log 2>&1; echo "driver rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2w_driver.json"))
for k in ("resident_gpu_prior",):
    e=d[k]; print(k, e["s_per_view"], {x:e[x] for x in ("wall_s","setup_s","fusion_s","fusion_kernel_ms","fusion_points","process_wall_s","fusion_phases")})
PY
