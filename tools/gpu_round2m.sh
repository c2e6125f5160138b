#!/bin/bash
# 2 GPUs: multi-device driver tests (park / shared pool per device), driver wall clock at --gpus 2, C3 at N=2
python -m pytest tests -m gpu -q -k "multi_gpu" 2>&1 | tail -8 > gpurun_out/r2m_pytest_n2.log; tail -4 gpurun_out/r2m_pytest_n2.log
python - <<'PY' > gpurun_out/r2m_driver_n2.log 2>&1
import json, subprocess, sys, tempfile, shutil
sys.path.insert(0, "acmmp-spherical_b200")
from acmmp_b200 import synth
scene = synth.make_pinhole_scene(n_views=12, width=3200, height=2130, focal=2800.0, seed=2)
tmp = tempfile.mkdtemp(prefix="drv_", dir="/dev/shm")
synth.write_dense_folder(scene, tmp, pgm=True)
for g in ("1", "2"):
    r = subprocess.run(["acmmp-spherical_b200/lib/acmmp_b200", tmp, "--seed", "11", "--resident", "1", "--gpu-prior", "1", "--gpus", g, "--fusion", "0"], capture_output=True, text=True)
    print("gpus", g, r.returncode, r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-800:])
shutil.rmtree(tmp, ignore_errors=True)
PY
cat gpurun_out/r2m_driver_n2.log | cut -c1-900
