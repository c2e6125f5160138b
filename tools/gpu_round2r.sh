#!/bin/bash
python -m pytest tests -m gpu -q -k "cpp_driver or reserved or parked" 2>&1 | tail -6 > gpurun_out/r2r_pytest.log; tail -3 gpurun_out/r2r_pytest.log
for rep in 1 2; do
python tools/driver_bench.py --views 11 --skip-files --no-fusion --out gpurun_out/r2r_driver_$rep.json > gpurun_out/r2r_driver_$rep.log 2>&1; echo "driver rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2r_driver_$rep.json"))
for k in ("resident","resident_gpu_prior"):
    e=d[k]; print(k, e["s_per_view"], {x:e[x] for x in ("setup_s","load_s","views_s","ctx_s","run_s","prior_dev_s","output_s","sweep1_s","geom_s","kernel_ms")})
PY
done
python bench.py --steps 3 --warmup 3 > gpurun_out/r2r_bench_c2.json 2> gpurun_out/r2r_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2r_bench_c2.json").read().replace("NaN","null"))
print({k:d.get(k) for k in ("value","ms_per_step")}, d["e2e"]["value"], d.get("e2e_driver"))
PY
