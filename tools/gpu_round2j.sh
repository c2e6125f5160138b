#!/bin/bash
# driver wall-clock breakdown (11 views C2 shape), C4 bench line with the issue-bound roofline, C3 at N=1
python tools/driver_bench.py --views 11 --skip-files --out gpurun_out/r2j_driver.json > gpurun_out/r2j_driver.log 2>&1; echo "driver rc=$?"; tail -5 gpurun_out/r2j_driver.log
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r2j_bench_c4.json 2> gpurun_out/r2j_bench_c4.err; echo "c4 rc=$?"
python bench.py --config C3 --steps 1 --warmup 3 > gpurun_out/r2j_bench_c3_n1.json 2> gpurun_out/r2j_bench_c3_n1.err; echo "c3 rc=$?"
tail -c 600 gpurun_out/r2j_bench_c3_n1.json
