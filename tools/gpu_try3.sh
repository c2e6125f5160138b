#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log | cut -c1-200
python -m pytest tests -m gpu -q -x -k "fusion or ncc_fixed or cpp_driver_runs" 2>&1 | tail -3
