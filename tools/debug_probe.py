"""Development aid: one probe launch on a small scene (used under compute-sanitizer)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "acmmp-spherical_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import util
from acmmp_b200 import Context
scene = util.pinhole_scene()
imgs, cams, ids = scene.problem(0)
ctx = Context(0)
ctx.set_views(imgs, cams)
planes = util.random_planes(scene, 0, seed=3, perturb=0.0)
out = ctx.probe_warp(planes, 1)
print("warp ok", float(np.nanmean(out[..., 0])))
c = ctx.probe_ncc(planes, 1)
print("ncc ok", float(c.mean()))
