"""Host-side planar-prior stage (reference ACMMP.cpp:904-1011 + main.cpp:113-185), CPU, OUT OF SCOPE as a
kernel target: it turns a photometric depth/cost map into the
(plane parameters, triangle mask) pair that `CudaPlanarPriorInitialization` consumes.

The reference uses cv::Subdiv2D for the Delaunay triangulation; Python's cv2.Subdiv2D is the same
implementation, so the triangulation is identical.  The rasteriser is cv2.fillConvexPoly instead of
the reference's barycentric stepping loop (main.cpp:153-159) -- coverage differs on triangle edges
only -- and the 3-point plane (cv::SVD::solveZ of a 3x4 system, ACMMP.cpp:956-989) is computed in
closed form (cross product), which spans the same null space.
"""
from __future__ import annotations

import cv2
import numpy as np

from . import MODEL_SPHERE


def support_points(costs, step=5):
    """GetSupportPoints, ACMMP.cpp:904-930: per 5x5 cell the pixel of least cost if that cost < 0.1."""
    H, W = costs.shape
    Hp, Wp = -(-H // step) * step, -(-W // step) * step
    c = np.full((Hp, Wp), np.inf, np.float32)
    c[:H, :W] = np.where(costs < 2.0, costs, np.inf)
    # reference scan order: col outer, row outer ... strict '>' keeps the FIRST minimum in (c, r) order
    cells = c.reshape(Hp // step, step, Wp // step, step).transpose(2, 0, 3, 1).reshape(Wp // step, Hp // step, step * step)
    idx = np.argmin(cells, axis=-1)               # index = dc * step + dr  (column-major inside the cell)
    best = np.take_along_axis(cells, idx[..., None], axis=-1)[..., 0]
    ok = best < 0.1
    cx, cy = np.nonzero(ok)
    xs = cx * step + idx[ok] // step
    ys = cy * step + idx[ok] % step
    return np.stack([xs, ys], axis=1).astype(np.int32)      # ordered col-cell major like the reference


def _pixel_dir(cam, xs, ys, W, H):
    if cam.model == MODEL_SPHERE:
        lon = (xs - cam.params[1]) / W * 2.0 * np.pi
        lat = -(ys - cam.params[2]) / H * np.pi
        return np.stack([np.cos(lat) * np.sin(lon), -np.sin(lat), np.cos(lat) * np.cos(lon)], axis=-1)
    return np.stack([(xs - cam.K[2]) / cam.K[0], (ys - cam.K[5]) / cam.K[4], np.ones_like(xs, dtype=np.float64)], axis=-1)


def planar_prior(cam, depths, costs, depth_min, depth_max):
    """Returns (plane_params [n,4] float32, masks [H,W] float32 with 1-based triangle ids)."""
    H, W = depths.shape
    pts = support_points(costs)
    masks = np.zeros((H, W), np.float32)
    if len(pts) < 3:
        return np.zeros((0, 4), np.float32), masks
    sub = cv2.Subdiv2D((0, 0, W, H))
    sub.insert([(float(x), float(y)) for x, y in pts])
    tri = sub.getTriangleList().astype(np.int32).reshape(-1, 3, 2)         # (int) truncation, ACMMP.cpp:948-950
    inside = np.all((tri[..., 0] >= 0) & (tri[..., 0] < W) & (tri[..., 1] >= 0) & (tri[..., 1] < H), axis=1)
    tri = tri[inside]
    n = len(tri)
    # GetPriorPlaneParams, ACMMP.cpp:956-989: plane through the three lifted support points
    xs, ys = tri[..., 0].astype(np.float64), tri[..., 1].astype(np.float64)
    d = depths[tri[..., 1], tri[..., 0]].astype(np.float64)
    P = _pixel_dir(cam, xs, ys, W, H) * d[..., None]                         # Get3DPointonRefCam, ACMMP.cpp:287-312
    nrm = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
    w = -(nrm * P[:, 0]).sum(-1)
    norm2 = np.linalg.norm(nrm, axis=-1)
    norm2 = np.where(w < 0, -norm2, norm2)
    norm2 = np.where(norm2 == 0, 1.0, norm2)
    params = np.concatenate([nrm / norm2[:, None], (w / norm2)[:, None]], axis=1).astype(np.float32)
    ids = np.zeros((H, W), np.int32)
    for i in range(n):
        cv2.fillConvexPoly(ids, tri[i], int(i + 1))
    # prior-depth validity (main.cpp:168-181): drop pixels whose prior depth leaves [depth_min, depth_max]
    ys_, xs_ = np.nonzero(ids)
    pp = params[ids[ys_, xs_] - 1].astype(np.float64)
    if cam.model == MODEL_SPHERE:
        dirs = _pixel_dir(cam, xs_.astype(np.float64), ys_.astype(np.float64), W, H)
        denom = (pp[:, :3] * dirs).sum(-1)
        dd = np.where(np.abs(denom) < 1e-6, 1e6, -pp[:, 3] / np.where(denom == 0, 1, denom))
    else:                                                                    # GetDepthFromPlaneParam, ACMMP.cpp:1009
        K = cam.K
        dd = -pp[:, 3] * K[0] / ((xs_ - K[2]) * pp[:, 0] + (K[0] / K[4]) * (ys_ - K[5]) * pp[:, 1] + K[0] * pp[:, 2])
    bad = ~((dd <= depth_max) & (dd >= depth_min))
    ids[ys_[bad], xs_[bad]] = 0
    masks[:] = ids
    return params, masks
