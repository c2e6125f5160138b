#!/bin/bash
# like gpu_try.sh plus the no-TMA debug path
tag=$1; v=$2
bash tools/gpu_try.sh $tag $v
echo "== no-TMA path"; ACMMP_B200_LIB=$PWD/acmmp-spherical_b200/lib/libacmmp_b200_$v.so ACMMP_NO_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "single_pass_photometric" > gpurun_out/pytest_notma.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_notma.log
