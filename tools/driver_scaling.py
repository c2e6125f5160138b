"""C++ driver on a ring scene at --gpus 1 and --gpus N (resident schedule, planar prior on the device): wall clock and the
per-device schedule time.  The views are rendered by a process pool.  usage: driver_scaling.py --views 32 --gpus 4 --out f.json"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import shutil
import time
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "acmmp-spherical_b200"))


def render(args):
    n_views, v, w, h, focal = args
    from acmmp_b200 import synth
    full = synth.make_pinhole_scene(n_views=n_views, width=w, height=h, focal=focal, seed=3, n_src=10, ring=True, render_ids=[])
    img, dep = synth._render(full.quads, synth.MODEL_PINHOLE, full.Rs[v], full.ts[v], w, h, K=full.Ks[v])
    return v, img, dep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=32)
    ap.add_argument("--gpus", type=int, default=4)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from acmmp_b200 import synth
    W, H, F = 3200, 2130, 2800.0
    t0 = time.time()
    scene = synth.make_pinhole_scene(n_views=a.views, width=W, height=H, focal=F, seed=3, n_src=10, ring=True, render_ids=[])
    with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        for v, img, dep in ex.map(render, [(a.views, v, W, H, F) for v in range(a.views)]):
            scene.images[v], scene.depths_gt[v] = img, dep
    tmp = tempfile.mkdtemp(prefix="drv_scale_", dir="/dev/shm")
    synth.write_dense_folder(scene, tmp, pgm=True)
    res = {"views": a.views, "scene_s": time.time() - t0}
    driver = str(ROOT / "acmmp-spherical_b200" / "lib" / "acmmp_b200")
    for g in (a.gpus, 1):
        r = subprocess.run([driver, tmp, "--seed", "11", "--resident", "1", "--gpu-prior", "1", "--gpus", str(g), "--fusion", "0"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1500:]
        d = json.loads(r.stdout.strip().splitlines()[-1])
        d["schedule_s"] = d["sweep1_s"] + d["geom_s"]
        res[f"gpus{g}"] = d
        print("gpus", g, {k: d.get(k) for k in ("wall_s", "setup_s", "sweep1_s", "geom_s", "schedule_s", "barrier_s", "exchange_s", "views_s", "run_s", "output_s", "kernel_ms")})
        trace = [l for l in r.stderr.splitlines() if l.startswith("[acmmp trace]")]
        if trace:
            d["trace"] = trace
            print("\n".join(l for l in trace if any(k in l for k in ("set_views ", "pool.", "park", "synchronize", "download_result", "export_depth", "set_depths", "prior_from_triangles ", "support_points", "next_level "))))
    res["schedule_speedup"] = res["gpus1"]["schedule_s"] / res[f"gpus{a.gpus}"]["schedule_s"]
    res["wall_speedup"] = res["gpus1"]["wall_s"] / res[f"gpus{a.gpus}"]["wall_s"]
    print("schedule speed-up", res["schedule_speedup"], "wall", res["wall_speedup"])
    shutil.rmtree(tmp, ignore_errors=True)
    if a.out:
        json.dump(res, open(a.out, "w"))


if __name__ == "__main__":
    main()
