#!/bin/bash
# device-wide buffer pool + acmmp_park: full GPU suite, then the driver's wall-clock breakdown with the library trace
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2l_pytest.log; tail -6 gpurun_out/r2l_pytest.log
ACMMP_TRACE=1 python tools/driver_bench.py --views 11 --skip-files --no-fusion --trace gpurun_out/r2l_trace --out gpurun_out/r2l_driver.json > gpurun_out/r2l_driver.log 2>&1; echo "driver rc=$?"; tail -3 gpurun_out/r2l_driver.log | cut -c1-1800
for f in gpurun_out/r2l_trace_*.txt; do echo "== $f"; cat $f; done
nvidia-smi --query-gpu=memory.used --format=csv
