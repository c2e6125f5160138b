#!/bin/bash
python -m pytest tests -m gpu -q -k "multi_gpu or resident_schedule" 2>&1 | tail -8 > gpurun_out/r2x_pytest_n2.log; tail -5 gpurun_out/r2x_pytest_n2.log | cut -c1-600
