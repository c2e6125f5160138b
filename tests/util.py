"""Shared helpers of the parity tests: scenes, comparison metrics, metric dumps."""
import json
import os
from functools import lru_cache
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"


def dump(name, d):
    """Write the measured parity figures next to the test result (gpurun_out/ travels back)."""
    try:
        OUT.mkdir(exist_ok=True)
        with open(OUT / f"metrics_{name}.json", "w") as f:
            json.dump(d, f, indent=1, default=float)
    except OSError:
        pass
    print(name, json.dumps(d, default=float))


@lru_cache(maxsize=None)
def pinhole_scene(width=320, height=240, focal=250.0, n_views=5, seed=1):
    from acmmp_b200 import synth
    return synth.make_pinhole_scene(n_views=n_views, width=width, height=height, focal=focal, seed=seed)


@lru_cache(maxsize=None)
def sphere_scene(width=512, height=256, n_views=5, seed=4):
    from acmmp_b200 import synth
    return synth.make_sphere_scene(n_views=n_views, width=width, height=height, seed=seed)


def scene_of(model):
    return pinhole_scene() if model == "pinhole" else sphere_scene()


def close_frac(a, b, atol, rtol=0.0, mask=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    ok |= (np.isnan(a) & np.isnan(b))
    if mask is not None:
        ok = ok[mask]
    return float(ok.mean()) if ok.size else 1.0


def interior(H, W, border):
    m = np.zeros((H, W), bool)
    m[border:H - border, border:W - border] = True
    return m


def colour_mask(H, W, colour):
    ys, xs = np.mgrid[0:H, 0:W]
    return ((xs + ys) & 1) == colour


def random_planes(scene, view, seed, perturb=0.0):
    """GT planes, optionally with the normal jittered and d scaled -- 'fixed plane hypotheses'."""
    from acmmp_b200 import synth
    rng = np.random.default_rng(seed)
    pl = synth.gt_planes(scene, view).copy()
    if perturb > 0:
        n = pl[..., :3] + perturb * rng.standard_normal(pl[..., :3].shape).astype(np.float32)
        n /= np.linalg.norm(n, axis=-1, keepdims=True)
        pl[..., :3] = n
        pl[..., 3] *= (1.0 + perturb * rng.uniform(-1, 1, pl[..., 3].shape)).astype(np.float32)
    return pl


def angle_deg(n1, n2):
    d = np.clip((n1 * n2).sum(-1) / (np.linalg.norm(n1, axis=-1) * np.linalg.norm(n2, axis=-1) + 1e-20), -1, 1)
    return np.degrees(np.arccos(d))


def grid_prior(scene, view, cell=16, seed=7, drop=0.25):
    """A synthetic stand-in for the CPU planar-prior stage: one plane per cell x cell block
    (the GT plane at the block centre, slightly perturbed), a fraction of blocks unlabelled."""
    from acmmp_b200 import synth
    rng = np.random.default_rng(seed)
    gt = synth.gt_planes(scene, view)
    H, W = gt.shape[:2]
    params, masks = [], np.zeros((H, W), np.float32)
    for by in range(0, H, cell):
        for bx in range(0, W, cell):
            if rng.random() < drop:
                continue
            cy, cx = min(by + cell // 2, H - 1), min(bx + cell // 2, W - 1)
            p = gt[cy, cx].copy()
            p[:3] += 0.02 * rng.standard_normal(3).astype(np.float32)
            p[:3] /= np.linalg.norm(p[:3])
            p[3] *= 1.0 + 0.01 * rng.uniform(-1, 1)
            params.append(p)
            masks[by:by + cell, bx:bx + cell] = len(params)
    return np.asarray(params, np.float32), masks


def world_normals(scene, view, planes_cam):
    R = np.asarray(scene.Rs[view], np.float64)
    n = planes_cam[..., :3].astype(np.float64) @ R        # R^T n  (row-vector form)
    return n.astype(np.float32)
