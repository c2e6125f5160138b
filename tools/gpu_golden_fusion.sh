#!/bin/bash
python tests/golden/make_fusion_golden.py gpurun_out/golden > gpurun_out/golden_fusion.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/golden_fusion.log | cut -c1-300
