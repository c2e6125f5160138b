"""Quick A/B timing of one photometric stage (RunPatchMatch, 3 iterations): reference kernels
(oracle/_ref) vs this library, same synthetic inputs, CUDA-event times.  Development aid; the
contract benchmark is bench.py."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "acmmp-spherical_b200"))
sys.path.insert(0, str(ROOT))

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--focal", type=float, default=500.0)
    ap.add_argument("--views", type=int, default=5)
    ap.add_argument("--model", default="pinhole")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--geom", action="store_true", help="also time a geometric-consistency stage (neighbour depth maps = ground truth)")
    ap.add_argument("--tap-prune", type=float, default=None, help="SPHERE: acmmp_set_sphere_tap_pruning threshold (0 = sample every tap)")
    a = ap.parse_args()
    from acmmp_b200 import Context, synth
    t0 = time.time()
    if a.model == "pinhole":
        scene = synth.make_pinhole_scene(n_views=a.views, width=a.width, height=a.height, focal=a.focal, seed=2)
    else:
        scene = synth.make_sphere_scene(n_views=a.views, width=a.width, height=a.height, seed=4)
    gen_s = time.time() - t0
    imgs, cams, ids = scene.problem(0)
    res = dict(width=a.width, height=a.height, views=a.views, model=a.model, scene_gen_s=gen_s)
    ctx = Context(0)
    ctx.set_views(imgs, cams)
    ctx.set_seed(1234)
    if a.tap_prune is not None:
        ctx.set_sphere_tap_pruning(a.tap_prune)
    for rep in range(2):
        t0 = time.time()
        ctx.run_patch_match()
        wall = time.time() - t0
        res[f"mine_run{rep}"] = dict(wall_ms=1e3 * wall, **ctx.timings())
    pa, ca = ctx.get_result()
    if a.geom:
        for rep in range(2):
            ctx.reset_modes()
            ctx.set_geom_consistency(False)
            ctx.set_depth_maps([None] + [scene.depths_gt[i] for i in ids[1:]])
            ctx.run_patch_match()
            res[f"geom_run{rep}"] = dict(**ctx.timings())
        res["geom_checksum"] = float(np.nansum(ctx.get_result()[0][..., 3], dtype=np.float64))
    gt = scene.depths_gt[0]
    res["mine_vs_gt_1pct"] = float((np.abs(pa[..., 3] - gt) / gt <= 0.01).mean())
    if not a.no_ref:
        from oracle.ref_driver import RefACMMP
        ref = RefACMMP(imgs, cams, seed=1234)
        t0 = time.time()
        ms = ref.run_patch_match()
        res["ref_run"] = dict(wall_ms=1e3 * (time.time() - t0), event_ms=ms)
        pb, cb = ref.get_result()
        res["ref_vs_gt_1pct"] = float((np.abs(pb[..., 3] - gt) / gt <= 0.01).mean())
        res["ref_init_ms"] = ref.launch_init()
        res["ref_pass_ms"] = [ref.launch_pass(c, i) for i in range(3) for c in (0, 1)]
        rel = np.abs(pa[..., 3] - pb[..., 3]) / np.maximum(np.abs(pb[..., 3]), 1e-9)
        res["mine_vs_ref_1pct"] = float((rel <= 0.01).mean())
    s = json.dumps(res, default=float)
    print(s)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
