#!/bin/bash
ACMMP_TRACE=1 python tools/driver_scaling.py --views 32 --gpus 4 --out gpurun_out/r2_driver_scaling_n4.json > gpurun_out/r2_driver_scaling_n4.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/r2_driver_scaling_n4.log | cut -c1-420
