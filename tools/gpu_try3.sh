#!/bin/bash
python -m pytest tests/test_gpu_cpp_driver.py tests/test_gpu_edge_cases.py -m gpu -q -k "resident_schedule or equirectangular or parked or jpeg or multi_gpu" 2>&1 | tail -40 > gpurun_out/try3.log; tail -40 gpurun_out/try3.log | cut -c1-300
