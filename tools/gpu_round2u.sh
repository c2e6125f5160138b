#!/bin/bash
# C++ driver at --gpus 1 / 2 on a 16-view C2-size scene (10 source views each), after the pool / reservation changes
python - <<'PY' > gpurun_out/r2u_driver_n2.log 2>&1
import json, subprocess, sys, tempfile, shutil, time
sys.path.insert(0, "acmmp-spherical_b200")
from acmmp_b200 import synth
scene = synth.make_pinhole_scene(n_views=16, width=3200, height=2130, focal=2800.0, seed=2)
tmp = tempfile.mkdtemp(prefix="drv_", dir="/dev/shm")
synth.write_dense_folder(scene, tmp, pgm=True)
for g in ("1", "2", "2", "1"):
    t0 = time.time()
    r = subprocess.run(["acmmp-spherical_b200/lib/acmmp_b200", tmp, "--seed", "11", "--resident", "1", "--gpu-prior", "1", "--gpus", g, "--fusion", "0"], capture_output=True, text=True)
    print("gpus", g, r.returncode, round(time.time() - t0, 2), r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-800:])
shutil.rmtree(tmp, ignore_errors=True)
PY
cat gpurun_out/r2u_driver_n2.log | cut -c1-1000
