#!/bin/bash
# long-triangle raster spread over the grid: parity of the device prior stage, then the driver's wall-clock breakdown with the library trace
python -m pytest tests -m gpu -q -k "prior or cpp_driver" 2>&1 | tail -12 > gpurun_out/r2k_pytest.log; tail -6 gpurun_out/r2k_pytest.log
ACMMP_TRACE=1 python tools/driver_bench.py --views 11 --skip-files --no-fusion --trace gpurun_out/r2k_trace --out gpurun_out/r2k_driver.json > gpurun_out/r2k_driver.log 2>&1; echo "driver rc=$?"; tail -3 gpurun_out/r2k_driver.log | cut -c1-1500
for f in gpurun_out/r2k_trace_*.txt; do echo "== $f"; cat $f; done
