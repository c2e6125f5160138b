"""Wall-clock comparison of the C++ driver's schedules on an ETH3D-size dense folder (C2 shape: 3200x2130 pinhole views,
three pyramid levels): file-chained (the reference's schedule), GPU-resident, GPU-resident + planar prior on the device.
Checks that the three write bit-identical maps.  Development / measurement aid:
    python tools/driver_bench.py [--views 6] [--out gpurun_out/driver_bench.json]"""
import argparse
import json
import shutil
import struct
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "acmmp-spherical_b200"))
DRIVER = ROOT / "acmmp-spherical_b200" / "lib" / "acmmp_b200"


def read_dmb(path):
    raw = open(path, "rb").read()
    t, h, w, nb = struct.unpack("<4i", raw[:16])
    return np.frombuffer(raw[16:], np.uint32).reshape(h, w, nb)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--width", type=int, default=3200)
    ap.add_argument("--height", type=int, default=2130)
    ap.add_argument("--focal", type=float, default=2800.0)
    ap.add_argument("--out", default="")
    ap.add_argument("--skip-files", action="store_true", help="only the resident schedules (no bit comparison)")
    ap.add_argument("--no-fusion", action="store_true", help="pass --fusion 0 to the driver")
    ap.add_argument("--trace", default="", help="prefix: the driver's stderr (ACMMP_TRACE=1 table) goes to <prefix>_<variant>.txt")
    a = ap.parse_args()
    from acmmp_b200 import synth
    scene = synth.make_pinhole_scene(n_views=a.views, width=a.width, height=a.height, focal=a.focal, seed=2)
    tmp = Path(tempfile.mkdtemp(prefix="acmmp_driver_bench_"))
    base = tmp / "files"
    base.mkdir()
    synth.write_dense_folder(scene, str(base), pgm=True)
    variants = {"files": ("0", "0"), "resident": ("1", "0"), "resident_gpu_prior": ("1", "1")}
    res = {"views": a.views, "width": a.width, "height": a.height, "src_views": len(scene.pairs[0][1])}
    folders = {}
    if a.skip_files:
        variants.pop("files")
    for name, (resident, gpu_prior) in variants.items():
        folders[name] = base if name == "files" else tmp / name
        if name != "files":
            shutil.copytree(base, folders[name])
        t0 = time.time()
        r = subprocess.run([str(DRIVER), str(folders[name]), "--seed", "11", "--resident", resident, "--gpu-prior", gpu_prior]
                           + (["--fusion", "0"] if a.no_fusion else []), capture_output=True, text=True)
        if a.trace:
            open(f"{a.trace}_{name}.txt", "w").write(r.stderr)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res[name] = json.loads(r.stdout.strip().splitlines()[-1])
        res[name]["fusion_phases"] = [l for l in r.stdout.splitlines() if l.startswith("[CUDA Fusion] read maps")]
        res[name]["process_wall_s"] = time.time() - t0
        res[name]["s_per_view"] = res[name]["wall_s"] / a.views
    for name in ("resident", "resident_gpu_prior"):
        if a.skip_files:
            break
        same = True
        for v in range(a.views):
            for dmb in ("depths.dmb", "depths_geom.dmb", "normals.dmb", "costs.dmb"):
                same &= bool(np.array_equal(read_dmb(folders["files"] / "ACMMP" / ("2333_%08d" % v) / dmb),
                                            read_dmb(folders[name] / "ACMMP" / ("2333_%08d" % v) / dmb)))
        res[name]["bit_identical_to_files"] = same
    shutil.rmtree(tmp, ignore_errors=True)
    s = json.dumps(res)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
