#!/bin/bash
python -m pytest tests/test_gpu_fusion.py tests/test_gpu_cpp_driver.py -m gpu -q -k "fusion or runs_the_reference_schedule" 2>&1 | tail -30 > gpurun_out/try3.log; tail -30 gpurun_out/try3.log | cut -c1-400
