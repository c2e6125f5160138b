#!/bin/bash
# round-2 records: full GPU suite, smoke, both arms of C2 and C4, driver bench
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2t_pytest_gpu.log; tail -4 gpurun_out/r2t_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2t_smoke.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2t_bench_c2_ref.json 2> gpurun_out/r2t_bench_c2_ref.err; echo "c2 ref rc=$?"
python bench.py --steps 3 --warmup 3 > gpurun_out/r2t_bench_c2.json 2> gpurun_out/r2t_bench_c2.err; echo "c2 rc=$?"
python bench.py --config C4 --impl reference --steps 3 --warmup 3 > gpurun_out/r2t_bench_c4_ref.json 2> gpurun_out/r2t_bench_c4_ref.err; echo "c4 ref rc=$?"
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r2t_bench_c4.json 2> gpurun_out/r2t_bench_c4.err; echo "c4 rc=$?"
python tools/driver_bench.py --views 11 --skip-files --no-fusion --out gpurun_out/r2t_driver.json > gpurun_out/r2t_driver.log 2>&1; echo "driver rc=$?"
python - <<'PY'
import json
for f in ("c2_ref","c2","c4_ref","c4"):
    try:
        d=json.loads(open("gpurun_out/r2t_bench_%s.json" % f).read().replace("NaN","null"))
        e=d.get("e2e_driver") or {}
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d.get("e2e",{}).get("value"), d.get("ms_per_checkerboard_pass"), (d.get("roofline") or {}).get("frac"), e.get("s_per_view"), e.get("s_per_view_after_cuda_startup"), e.get("breakdown_s"))
    except Exception as ex: print(f, "unreadable", ex)
d=json.load(open("gpurun_out/r2t_driver.json"))
for k in ("resident","resident_gpu_prior"):
    e=d[k]; print(k, e["s_per_view"], {x:e[x] for x in ("setup_s","load_s","views_s","run_s","prior_dev_s","output_s","sweep1_s","geom_s","kernel_ms")})
PY
