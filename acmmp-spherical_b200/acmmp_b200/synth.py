"""Deterministic synthetic multi-view scenes (SURVEY.md section 8(d) "Synthetic inputs").

Textured planar quads rendered by ray casting into PINHOLE or equirectangular (SPHERE) cameras,
using exactly the fork's camera conventions (reference ACMMP.cpp:247-350): R, t map world to camera;
PINHOLE pixel = K X / z; SPHERE lon = (x-cx)/W*2pi, lat = -(y-cy)/H*pi, dir = (cos lat sin lon,
-sin lat, cos lat cos lon).  Returns float32 grey images (0..255, 8-bit quantised like a decoded
JPEG would be), `Camera` structs, ground-truth depth (z-depth for PINHOLE, radial for SPHERE --
what the reference's depth maps converge to) and the pair list.  Also writes the on-disk layout
(cams/, images/, pair.txt) in the formats the reference READS (ACMMP.cpp:146-209, main.cpp:4-33).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import cv2
import numpy as np

from . import MODEL_PINHOLE, MODEL_SPHERE, make_camera


@dataclass
class Quad:
    origin: np.ndarray      # world corner
    u: np.ndarray           # unit axis
    v: np.ndarray           # unit axis
    lu: float
    lv: float
    texture: np.ndarray     # float32 (rows along v, cols along u)
    texel: float            # world units per texel

    @property
    def normal(self):
        return np.cross(self.u, self.v)


@dataclass
class Scene:
    model: int
    images: list
    cams: list
    depths_gt: list
    pairs: list                  # [(ref_id, [src ids...])]
    Rs: list = field(default_factory=list)
    ts: list = field(default_factory=list)
    Ks: list = field(default_factory=list)
    quads: list = field(default_factory=list)

    def problem(self, ref):
        ids = [ref] + list(self.pairs[ref][1])
        return [self.images[i] for i in ids], [self.cams[i] for i in ids], ids


def _noise_texture(rng, rows, cols, sigma=1.5):
    t = rng.random((rows, cols), dtype=np.float32)
    t = cv2.GaussianBlur(t, (0, 0), sigma)
    # second, coarser octave so that coarse pyramid levels still see texture
    c = rng.random((rows // 8 + 2, cols // 8 + 2), dtype=np.float32)
    c = cv2.GaussianBlur(c, (0, 0), 1.0)
    c = cv2.resize(c, (cols, rows), interpolation=cv2.INTER_CUBIC)
    t = (t - t.mean()) / (t.std() + 1e-9) + 0.8 * (c - c.mean()) / (c.std() + 1e-9)
    t = (t - t.min()) / (t.max() - t.min())
    return (20.0 + 215.0 * t).astype(np.float32)


def _make_quad(rng, origin, u, v, lu, lv, texel):
    u = np.asarray(u, np.float64); u /= np.linalg.norm(u)
    v = np.asarray(v, np.float64); v /= np.linalg.norm(v)
    rows, cols = int(np.ceil(lv / texel)) + 2, int(np.ceil(lu / texel)) + 2
    return Quad(np.asarray(origin, np.float64), u, v, float(lu), float(lv), _noise_texture(rng, rows, cols), float(texel))


def _look_at(C, target, up=(0.0, -1.0, 0.0)):
    """World -> camera rotation with +z forward, +x right, +y down."""
    f = np.asarray(target, np.float64) - np.asarray(C, np.float64)
    f /= np.linalg.norm(f)
    upv = np.asarray(up, np.float64)
    r = np.cross(-upv, f)      # right
    r /= np.linalg.norm(r)
    d = np.cross(f, r)         # down
    R = np.stack([r, d, f], axis=0)
    t = -R @ np.asarray(C, np.float64)
    return R, t


def _render(quads, model, R, t, width, height, K=None, cx=None, cy=None, chunk_rows=256):
    """Ray cast; returns (image float32, depth float32: z for PINHOLE, radial for SPHERE)."""
    C = -R.T @ t
    img = np.zeros((height, width), np.float32)
    depth = np.zeros((height, width), np.float32)
    for y0 in range(0, height, chunk_rows):
        y1 = min(height, y0 + chunk_rows)
        ys, xs = np.mgrid[y0:y1, 0:width].astype(np.float64)
        if model == MODEL_PINHOLE:
            dc = np.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], np.ones_like(xs)], axis=-1)
        else:
            lon = (xs - cx) / width * 2.0 * np.pi
            lat = -(ys - cy) / height * np.pi
            dc = np.stack([np.cos(lat) * np.sin(lon), -np.sin(lat), np.cos(lat) * np.cos(lon)], axis=-1)
        dw = dc @ R            # world direction = R^T d_cam
        best_t = np.full(xs.shape, np.inf)
        val = np.zeros(xs.shape, np.float32)
        for q in quads:
            n = q.normal
            denom = dw @ n
            with np.errstate(divide="ignore", invalid="ignore"):
                tt = ((q.origin - C) @ n) / denom
            P = C + dw * tt[..., None]
            a = (P - q.origin) @ q.u
            b = (P - q.origin) @ q.v
            hit = (tt > 1e-6) & (tt < best_t) & (a >= 0) & (a <= q.lu) & (b >= 0) & (b <= q.lv) & np.isfinite(tt)
            if not hit.any():
                continue
            mapx = (a / q.texel).astype(np.float32)
            mapy = (b / q.texel).astype(np.float32)
            mapx[~hit] = 0
            mapy[~hit] = 0
            s = cv2.remap(q.texture, mapx, mapy, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
            val = np.where(hit, s, val)
            best_t = np.where(hit, tt, best_t)
        finite = np.isfinite(best_t)
        img[y0:y1] = np.where(finite, val, 0.0)
        # parametrisation along dc: PINHOLE dc.z == 1 -> z-depth; SPHERE |dc| == 1 -> radial
        depth[y0:y1] = np.where(finite, best_t, 0.0).astype(np.float32)
    img = np.clip(np.rint(img), 0, 255).astype(np.float32)      # 8-bit quantisation
    return img, depth


def make_pinhole_scene(n_views=5, width=640, height=480, focal=500.0, seed=1, n_src=None, depth0=3.0,
                       baseline_ratio=0.07, ring=False, render_ids=None) -> Scene:
    """Background plane + three tilted foreground quads, cameras on an arc (or ring) around them.
    render_ids: render only these reference views and their sources (the others get None images / depths): a rank of a multi-GPU
    job only needs its own reference view and that view's sources."""
    rng = np.random.default_rng(seed)
    texel = 0.7 * depth0 / focal
    # arc: the background grows with the camera spread; ring: the cameras stay inside a fixed circle (a 64-view arc would
    # need a background texture wider than cv2.remap's 32767-texel limit)
    halfw = 0.5 * width / focal * depth0 * 1.9 + (0.0 if ring else baseline_ratio * depth0 * n_views)
    halfh = 0.5 * height / focal * depth0 * 1.9
    quads = [_make_quad(rng, (-halfw, -halfh, depth0 * 1.15), (1, 0, 0), (0, 1, 0), 2 * halfw, 2 * halfh, texel)]
    fw, fh = 0.45 * halfw, 0.5 * halfh
    quads.append(_make_quad(rng, (-0.8 * halfw * 0.6, -0.6 * halfh, depth0 * 0.95), (1, 0, 0.25), (0, 1, 0.0), fw, fh, texel))
    quads.append(_make_quad(rng, (0.05 * halfw, -0.1 * halfh, depth0 * 1.02), (1, 0.0, -0.2), (0, 1, 0.15), fw, fh, texel))
    quads.append(_make_quad(rng, (-0.3 * halfw, 0.15 * halfh, depth0 * 0.85), (1, 0.1, 0.0), (-0.1, 1, 0.3), 0.8 * fw, 0.7 * fh, texel))
    K = np.array([[focal, 0, width / 2.0], [0, focal, height / 2.0], [0, 0, 1]], np.float64)
    step = baseline_ratio * depth0
    Rs, ts, images, depths, cams = [], [], [], [], []
    for i in range(n_views):
        if ring:
            ang = 2 * np.pi * i / n_views
            C = np.array([0.35 * halfw * np.cos(ang), 0.35 * halfh * np.sin(ang), 0.04 * depth0 * np.sin(2 * ang)])
        else:
            off = (i - (n_views - 1) / 2.0) * step
            C = np.array([off, 0.15 * step * ((i % 3) - 1), 0.05 * step * ((i % 2) * 2 - 1)])
        R, t = _look_at(C, (0.15 * C[0], 0.1 * C[1], depth0))
        Rs.append(R); ts.append(t)
    pairs = _nearest_pairs(Rs, ts, n_views, n_src if n_src is not None else n_views - 1)
    if render_ids is not None:
        render_ids = set(render_ids)
        for r in list(render_ids):
            render_ids.update(pairs[r][1])
    for i in range(n_views):
        if render_ids is None or i in render_ids:
            img, dep = _render(quads, MODEL_PINHOLE, Rs[i], ts[i], width, height, K=K)
        else:
            img, dep = None, None
        images.append(img); depths.append(dep)
    if render_ids is None:
        dmin = min(float(d[d > 0].min()) for d in depths) * 0.9
        dmax = max(float(d.max()) for d in depths) * 1.1
    else:       # the same range on every rank: from the scene geometry, not from the rendered subset
        dmin, dmax = 0.6 * depth0, 1.5 * depth0
    for i in range(n_views):
        cams.append(make_camera(MODEL_PINHOLE, Rs[i], ts[i], K=K, width=width, height=height, depth_min=dmin, depth_max=dmax))
    return Scene(MODEL_PINHOLE, images, cams, depths, pairs, Rs, ts, [K] * n_views, quads)


def make_sphere_scene(n_views=5, width=1024, height=512, seed=4, n_src=None, room=(8.0, 5.0, 6.0), spread=0.5,
                      render_ids=None) -> Scene:
    """Textured box room seen by equirectangular cameras near its centre.
    render_ids: as in make_pinhole_scene (only these reference views and their sources are rendered)."""
    rng = np.random.default_rng(seed)
    lx, ly, lz = room
    texel = 0.7 * (2 * np.pi * 0.5 * min(room)) / width * 0.5
    hx, hy, hz = lx / 2, ly / 2, lz / 2
    quads = [
        _make_quad(rng, (-hx, -hy, hz), (1, 0, 0), (0, 1, 0), lx, ly, texel),      # front  z=+hz
        _make_quad(rng, (hx, -hy, -hz), (-1, 0, 0), (0, 1, 0), lx, ly, texel),     # back   z=-hz
        _make_quad(rng, (-hx, -hy, -hz), (0, 0, 1), (0, 1, 0), lz, ly, texel),     # left   x=-hx
        _make_quad(rng, (hx, -hy, hz), (0, 0, -1), (0, 1, 0), lz, ly, texel),      # right  x=+hx
        _make_quad(rng, (-hx, -hy, -hz), (1, 0, 0), (0, 0, 1), lx, lz, texel),     # ceiling y=-hy
        _make_quad(rng, (-hx, hy, hz), (1, 0, 0), (0, 0, -1), lx, lz, texel),      # floor  y=+hy
    ]
    Rs, ts, images, depths, cams = [], [], [], [], []
    cx, cy = width / 2.0, height / 2.0
    for i in range(n_views):
        ang = 2 * np.pi * i / max(n_views, 1)
        C = np.array([spread * np.cos(ang), 0.15 * spread * np.sin(2 * ang), spread * np.sin(ang)]) if i > 0 else np.zeros(3)
        yaw = 0.2 * np.sin(ang)
        R = np.array([[np.cos(yaw), 0, -np.sin(yaw)], [0, 1, 0], [np.sin(yaw), 0, np.cos(yaw)]], np.float64)
        t = -R @ C
        Rs.append(R); ts.append(t)
    pairs = _nearest_pairs(Rs, ts, n_views, n_src if n_src is not None else n_views - 1)
    wanted = None
    if render_ids is not None:
        wanted = set(render_ids)
        for r in list(wanted):
            wanted.update(pairs[r][1])
    for i in range(n_views):
        if wanted is None or i in wanted:
            img, dep = _render(quads, MODEL_SPHERE, Rs[i], ts[i], width, height, cx=cx, cy=cy)
        else:
            img, dep = None, None
        images.append(img); depths.append(dep)
    if wanted is None:
        dmin = min(float(d[d > 0].min()) for d in depths) * 0.9
        dmax = max(float(d.max()) for d in depths) * 1.1
    else:       # the same range on every rank: from the room, not from the rendered subset
        dmin = 0.9 * (0.5 * min(room) - 1.2 * spread)
        dmax = 1.1 * (0.5 * float(np.linalg.norm(room)) + 1.2 * spread)
    for i in range(n_views):
        cams.append(make_camera(MODEL_SPHERE, Rs[i], ts[i], sphere=(1.0, cx, cy, 0.0), width=width, height=height,
                                depth_min=dmin, depth_max=dmax))
    return Scene(MODEL_SPHERE, images, cams, depths, pairs, Rs, ts, [None] * n_views, quads)


def _nearest_pairs(Rs, ts, n, n_src):
    Cs = [-(R.T @ t) for R, t in zip(Rs, ts)]
    pairs = []
    for i in range(n):
        d = sorted((float(np.linalg.norm(Cs[i] - Cs[j])), j) for j in range(n) if j != i)
        pairs.append((i, [j for _, j in d[:n_src]]))
    return pairs


def scale_problem(images, cams, max_size):
    """What InuputInitialization does per image (reference ACMMP.cpp:605-643): cv::resize INTER_LINEAR
    to round(size * factor) when an image exceeds max_size, and scale K / (cx, cy)."""
    from . import Camera
    out_imgs, out_cams = [], []
    for im, cam in zip(images, cams):
        rows, cols = im.shape
        c = Camera.from_buffer_copy(cam)
        if cols <= max_size and rows <= max_size:
            out_imgs.append(im); c.width, c.height = cols, rows; out_cams.append(c)
            continue
        factor = min(np.float32(max_size) / np.float32(cols), np.float32(max_size) / np.float32(rows))
        # std::round (ACMMP.cpp:620-621) rounds halves away from zero -- 2130 / 4 = 532.5 -> 533; Python's round() gives 532
        new_cols = int(np.floor(float(np.float32(cols) * factor) + 0.5))
        new_rows = int(np.floor(float(np.float32(rows) * factor) + 0.5))
        sx = np.float32(new_cols) / np.float32(cols)
        sy = np.float32(new_rows) / np.float32(rows)
        scaled = cv2.resize(im, (new_cols, new_rows), interpolation=cv2.INTER_LINEAR)
        if c.model == MODEL_SPHERE:
            c.params[1] = float(np.float32(c.params[1]) * sx)
            c.params[2] = float(np.float32(c.params[2]) * sy)
        else:
            c.K[0] = float(np.float32(c.K[0]) * sx); c.K[2] = float(np.float32(c.K[2]) * sx)
            c.K[4] = float(np.float32(c.K[4]) * sy); c.K[5] = float(np.float32(c.K[5]) * sy)
        c.width, c.height = new_cols, new_rows
        out_imgs.append(np.ascontiguousarray(scaled, np.float32)); out_cams.append(c)
    return out_imgs, out_cams


def gt_planes(scene: Scene, view: int):
    """Per-pixel (camera-frame normal, d) of the true surface in the fork's plane convention:
    n . (t * unit_ray) + d = 0 with t the depth the reference stores."""
    cam = scene.cams[view]
    H, W = scene.images[view].shape
    R, t = scene.Rs[view], scene.ts[view]
    C = -R.T @ t
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    if scene.model == MODEL_PINHOLE:
        K = scene.Ks[view]
        dc = np.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], np.ones_like(xs)], axis=-1)
    else:
        lon = (xs - cam.params[1]) / W * 2.0 * np.pi
        lat = -(ys - cam.params[2]) / H * np.pi
        dc = np.stack([np.cos(lat) * np.sin(lon), -np.sin(lat), np.cos(lat) * np.cos(lon)], axis=-1)
    unit = dc / np.linalg.norm(dc, axis=-1, keepdims=True)
    depth = scene.depths_gt[view].astype(np.float64)
    planes = np.zeros((H, W, 4), np.float32)
    # find the quad hit at each pixel from the depth: nearest plane to the lifted point
    P = C + (dc @ R) * depth[..., None]
    best = np.full((H, W), np.inf)
    for q in scene.quads:
        dist = np.abs((P - q.origin) @ q.normal)
        n_cam = R @ q.normal
        n_cam = np.broadcast_to(n_cam, (H, W, 3)).copy()
        flip = (n_cam * unit).sum(-1) > 0
        n_cam[flip] *= -1
        sel = dist < best
        best = np.where(sel, dist, best)
        d = -(n_cam * unit).sum(-1) * depth
        planes[sel, :3] = n_cam[sel].astype(np.float32)
        planes[sel, 3] = d[sel].astype(np.float32)
    return planes


def write_dense_folder(scene: Scene, folder: str, jpeg_quality=95, pgm=False):
    """cams/%08d_cam.txt, images/%08d.jpg, pair.txt in the formats the reference reads.
    pgm=True also writes lossless images/%08d.pgm twins (the C++ host prefers them: identical pixels
    for every decoder)."""
    os.makedirs(os.path.join(folder, "cams"), exist_ok=True)
    os.makedirs(os.path.join(folder, "images"), exist_ok=True)
    for i, (img, cam) in enumerate(zip(scene.images, scene.cams)):
        cv2.imwrite(os.path.join(folder, "images", "%08d.jpg" % i), img.astype(np.uint8), [cv2.IMWRITE_JPEG_QUALITY, jpeg_quality])
        if pgm:
            cv2.imwrite(os.path.join(folder, "images", "%08d.pgm" % i), img.astype(np.uint8))
        with open(os.path.join(folder, "cams", "%08d_cam.txt" % i), "w") as f:
            f.write("extrinsic\n")
            for r in range(3):
                f.write("%.9g %.9g %.9g %.9g\n" % (cam.R[3 * r], cam.R[3 * r + 1], cam.R[3 * r + 2], cam.t[r]))
            f.write("0.0 0.0 0.0 1.0\n\nintrinsic\n")
            if cam.model == MODEL_SPHERE:
                f.write("SPHERE\n%.9g %.9g %.9g\n\n" % (cam.params[0], cam.params[1], cam.params[2]))
                # SPHERE branch reads: depth_min depth_interval n_planes depth_max (ACMMP.cpp:185-192)
                f.write("%.9g %.9g %d %.9g\n" % (cam.depth_min, (cam.depth_max - cam.depth_min) / 191.0, 192, cam.depth_max))
            else:
                for r in range(3):
                    f.write("%.9g %.9g %.9g\n" % (cam.K[3 * r], cam.K[3 * r + 1], cam.K[3 * r + 2]))
                # PINHOLE branch reads: depth_min depth_max _ _ (ACMMP.cpp:204-205)
                f.write("\n%.9g %.9g 0 0\n" % (cam.depth_min, cam.depth_max))
    with open(os.path.join(folder, "pair.txt"), "w") as f:
        f.write("%d\n" % len(scene.pairs))
        for ref, srcs in scene.pairs:
            f.write("%d\n%d " % (ref, len(srcs)))
            f.write(" ".join("%d %.3f" % (s, 1.0) for s in srcs) + "\n")
