"""COLMAP sparse model -> the dense-folder contract the PatchMatch path reads.

Replaces the fork's offline converter (reference colmap2mvsnet_acm.py:249-406): `cams/%08d_cam.txt`, `pair.txt`,
`images/%08d.jpg` from `<dense>/sparse/{cameras,images,points3D}.{txt,bin}` + `<dense>/images/`.  Same outputs for the
same model (tests/test_cpu_colmap.py compares with files the reference script wrote for a committed fixture):

  * images are re-indexed 0..N-1 in the order of their COLMAP ids (colmap2mvsnet_acm.py:262);
  * depth range per image from its own 3-D points: sorted depths (z for pinhole models, radius for SPHERE), the values at
    20 % and 80 % scaled by 0.75 / 1.25; `depth_num` planes (--max_d, or from a one-pixel baseline when 0) (:183-217);
  * neighbour candidates: the top_k nearest camera centres (KD-tree), pruned greedily by shared track count (>= min_shared,
    at most top_k per image), scored by the number of shared points when the 75th percentile of their triangulation
    angles reaches theta0 degrees, else 0 (:222-244, :304-356);
  * `*_cam.txt`: `extrinsic` + 4x4 row-major [R|t; 0 0 0 1], `intrinsic` + either `SPHERE\\n f cx cy` or the 3x3 K, then
    `dmin dinterval ndepth dmax` (:365-388); `pair.txt`: N, then per image `id\\n n  id score ...` (:391-397).

Structure is this repo's own: array-based model tables, no process pool (the scoring is a few vectorised operations per
pair), numbers written with repr() like the reference's str() of Python / numpy floats.
"""
from __future__ import annotations

import os
import shutil
import struct
from dataclasses import dataclass

import numpy as np

# COLMAP camera models: id -> (name, number of parameters); 11 = the fork's equirectangular model
_MODELS = {0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5), 4: ("OPENCV", 8),
           5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5), 8: ("SIMPLE_RADIAL_FISHEYE", 4),
           9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12), 11: ("SPHERE", 3)}
# which parameters are (fx, fy, cx, cy): models with a single focal length list it first
_SINGLE_F = {"SIMPLE_PINHOLE", "SIMPLE_RADIAL", "RADIAL", "SPHERE"}
_KNOWN = {"SIMPLE_PINHOLE", "PINHOLE", "SIMPLE_RADIAL", "RADIAL", "OPENCV", "OPENCV_FISHEYE", "FULL_OPENCV", "FOV",
          "THIN_PRISM_FISHEYE", "SPHERE"}


@dataclass
class ColmapCamera:
    id: int
    model: str
    width: int
    height: int
    params: np.ndarray


@dataclass
class ColmapImage:
    id: int
    qvec: np.ndarray          # (w, x, y, z)
    tvec: np.ndarray
    camera_id: int
    name: str
    point3D_ids: np.ndarray   # one per 2-D observation, -1 = no 3-D point


def _text_rows(path):
    with open(path) as f:
        for line in f:
            if line.strip() and not line.lstrip().startswith("#"):
                yield line.split()


def read_sparse_model(sparse_dir, ext=".txt"):
    """-> (cameras {id: ColmapCamera}, images {id: ColmapImage}, points {id: xyz float64[3]})"""
    cams, imgs, pts = {}, {}, {}
    if ext == ".txt":
        for s in _text_rows(os.path.join(sparse_dir, "cameras.txt")):
            cams[int(s[0])] = ColmapCamera(int(s[0]), s[1], int(s[2]), int(s[3]), np.array([float(v) for v in s[4:]]))
        with open(os.path.join(sparse_dir, "images.txt")) as f:
            lines = [l for l in f if not l.lstrip().startswith("#")]
        k = 0
        while k < len(lines):
            if not lines[k].strip():
                k += 1
                continue
            s = lines[k].split()
            track = lines[k + 1].split() if k + 1 < len(lines) else []
            imgs[int(s[0])] = ColmapImage(int(s[0]), np.array([float(v) for v in s[1:5]]), np.array([float(v) for v in s[5:8]]),
                                          int(s[8]), s[9], np.array([int(v) for v in track[2::3]], dtype=np.int64))
            k += 2
        for s in _text_rows(os.path.join(sparse_dir, "points3D.txt")):
            pts[int(s[0])] = np.array([float(v) for v in s[1:4]])
    else:
        def rd(f, fmt):
            return struct.unpack("<" + fmt, f.read(struct.calcsize("<" + fmt)))
        with open(os.path.join(sparse_dir, "cameras.bin"), "rb") as f:
            for _ in range(rd(f, "Q")[0]):
                cid, mid, w, h = rd(f, "iiQQ")
                name, npar = _MODELS[mid]
                cams[cid] = ColmapCamera(cid, name, w, h, np.array(rd(f, "d" * npar)))
        with open(os.path.join(sparse_dir, "images.bin"), "rb") as f:
            for _ in range(rd(f, "Q")[0]):
                vals = rd(f, "idddddddi")
                name = bytearray()
                while True:
                    c = f.read(1)
                    if c == b"\x00":
                        break
                    name += c
                n2d = rd(f, "Q")[0]
                obs = np.frombuffer(f.read(24 * n2d), dtype=np.dtype([("x", "<f8"), ("y", "<f8"), ("p", "<i8")]))
                imgs[vals[0]] = ColmapImage(vals[0], np.array(vals[1:5]), np.array(vals[5:8]), vals[8], name.decode(),
                                            obs["p"].astype(np.int64))
        with open(os.path.join(sparse_dir, "points3D.bin"), "rb") as f:
            for _ in range(rd(f, "Q")[0]):
                pid, x, y, z = rd(f, "Qddd")
                f.read(3 + 8)                          # rgb, error
                f.read(8 * rd(f, "Q")[0])              # track
                pts[pid] = np.array([x, y, z])
    return cams, imgs, pts


def rotation_of(qvec):
    """Unit quaternion (w, x, y, z) -> rotation matrix (COLMAP convention)."""
    w, x, y, z = qvec
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                     [2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


def intrinsic_matrix(cam: ColmapCamera):
    if cam.model not in _KNOWN:
        raise ValueError(f"camera model {cam.model} is not supported by the converter")
    p = cam.params
    fx, fy, cx, cy = (p[0], p[0], p[1], p[2]) if cam.model in _SINGLE_F else (p[0], p[1], p[2], p[3])
    K = np.eye(3)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx, fy, cx, cy
    return K


def depth_range(img: ColmapImage, cam: ColmapCamera, E, K, points, max_d, interval_scale):
    """(dmin, dinterval, depth_num, dmax) of one image, colmap2mvsnet_acm.py:183-217."""
    ids = img.point3D_ids[img.point3D_ids >= 0]
    X = np.stack([points[int(p)] for p in ids]) if len(ids) else np.zeros((0, 3))
    Xc = np.concatenate([X, np.ones((len(X), 1))], axis=1) @ E.T
    d = np.linalg.norm(Xc[:, :3], axis=1) if cam.model == "SPHERE" else Xc[:, 2]
    d = np.sort(d[d > 0])
    dmin = d[int(len(d) * 0.2)] * 0.75
    dmax = d[int(len(d) * 0.8)] * 1.25
    if max_d == 0:
        # number of planes such that neighbouring planes are one pixel apart at dmin
        p1 = np.array([K[0, 2], K[1, 2], 1.0])
        Rinv, Kinv = np.linalg.inv(E[:3, :3]), np.linalg.inv(K)
        P1 = Rinv @ ((Kinv @ p1) * dmin - E[:3, 3])
        P2 = Rinv @ ((Kinv @ (p1 + np.array([1.0, 0.0, 0.0]))) * dmin - E[:3, 3])
        depth_num = int((1 / dmin - 1 / dmax) / (1 / dmin - 1 / (dmin + np.linalg.norm(P2 - P1))))
    else:
        depth_num = max_d
    return dmin, (dmax - dmin) / (depth_num - 1) / interval_scale, depth_num, dmax


def convert(dense_folder, save_folder, model_ext=".txt", max_d=192, interval_scale=1.0, theta0=1.0, top_k=20, min_shared=10,
            copy_images=True):
    """Writes cams/, pair.txt (and images/) under save_folder; returns (N, view selection [[(neighbour, score), ...]])."""
    from scipy.spatial import cKDTree
    cams, imgs_raw, points = read_sparse_model(os.path.join(dense_folder, "sparse"), model_ext)
    order = sorted(imgs_raw)                                   # position i (0-based) <- i-th smallest COLMAP image id
    imgs = [imgs_raw[k] for k in order]
    N = len(imgs)
    os.makedirs(os.path.join(save_folder, "cams"), exist_ok=True)
    E = []
    for im in imgs:
        e = np.eye(4)
        e[:3, :3] = rotation_of(im.qvec)
        e[:3, 3] = im.tvec
        E.append(e)
    K = {cid: intrinsic_matrix(c) for cid, c in cams.items()}
    ranges = [depth_range(im, cams[im.camera_id], E[i], K[im.camera_id], points, max_d, interval_scale) for i, im in enumerate(imgs)]
    print("depth_ranges[1]", ranges[0] if ranges else None)
    centres = np.stack([-(e[:3, :3].T @ e[:3, 3]) for e in E])

    # candidate pairs: every image with its top_k nearest camera centres.  The pairs go through a Python set and a list
    # like in the reference (:318-330): the greedy pruning below walks them in descending shared count and is stable, so
    # the order of equal counts -- the set's iteration order -- decides which pairs survive.
    _, nn = cKDTree(centres).query(centres, k=top_k + 1)
    nn = nn.reshape(N, -1)
    candidates = set()
    for a in range(N):
        for b in nn[a]:
            if b == a or b >= N:
                continue
            candidates.add((min(a, int(b)), max(a, int(b))))
    pairs = list(candidates)
    tracks = [set(im.point3D_ids.tolist()) for im in imgs]      # -1 included, like the reference's set() of the id array
    shared = [len(tracks[a] & tracks[b]) for a, b in pairs]
    taken = [0] * N
    kept = []
    for (a, b), s in sorted(zip(pairs, shared), key=lambda t: t[1], reverse=True):
        if s < min_shared:
            break
        if taken[a] < top_k and taken[b] < top_k:
            taken[a] += 1
            taken[b] += 1
            kept.append((a, b))
    print(f"[INFO] Kept {len(kept)} pairs (<={top_k} per image, >={min_shared} shared tracks)")

    # score = number of shared tracks if enough of them are seen under a triangulation angle >= theta0 degrees
    score = np.zeros((N, N))
    for a, b in kept:
        common = tracks[a] & tracks[b]
        if not common:
            continue
        ids = [p for p in common if p != -1]
        if not ids:
            continue
        P = np.stack([points[p] for p in ids])
        va, vb = centres[a] - P, centres[b] - P
        cosang = np.clip((va * vb).sum(1) / (np.linalg.norm(va, axis=1) * np.linalg.norm(vb, axis=1)), -1.0, 1.0)
        if np.percentile(np.degrees(np.arccos(cosang)), 75) >= theta0:
            score[a, b] = score[b, a] = float(len(common))
    view_sel = []
    for i in range(N):
        best = np.argsort(score[i])[::-1]
        view_sel.append([(int(k), float(score[i, k])) for k in best if score[i, k] > 0][:top_k])

    for i, im in enumerate(imgs):
        cam = cams[im.camera_id]
        with open(os.path.join(save_folder, "cams", "%08d_cam.txt" % i), "w") as f:
            f.write("extrinsic\n")
            for r in range(4):
                f.write(" ".join(str(v) for v in E[i][r]) + "\n")
            f.write("\nintrinsic\n")
            if cam.model == "SPHERE":
                f.write("SPHERE\n")
                f.write(f"{cam.params[0]} {cam.params[1]} {cam.params[2]}\n")
            else:
                for r in range(3):
                    f.write(" ".join(str(v) for v in K[cam.id][r]) + "\n")
            d0, dint, nd, dmax = ranges[i]
            f.write(f"\n{d0} {dint} {nd} {dmax}\n")
    with open(os.path.join(save_folder, "pair.txt"), "w") as f:
        f.write(f"{N}\n")
        for i, nbrs in enumerate(view_sel):
            f.write(f"{i}\n{len(nbrs)} ")
            for j, s in nbrs:
                f.write(f"{j} {int(s)} ")
            f.write("\n")
    if copy_images:
        import cv2
        os.makedirs(os.path.join(save_folder, "images"), exist_ok=True)
        for i, im in enumerate(imgs):
            src = os.path.join(dense_folder, "images", im.name)
            dst = os.path.join(save_folder, "images", "%08d.jpg" % i)
            if src.lower().endswith(".jpg"):
                shutil.copyfile(src, dst)
            else:
                cv2.imwrite(dst, cv2.imread(src))
    return N, view_sel


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="COLMAP sparse model -> cams/, images/, pair.txt of the ACMMP dense-folder contract")
    ap.add_argument("--dense_folder", required=True, help="folder with sparse/ and images/")
    ap.add_argument("--save_folder", required=True)
    ap.add_argument("--model_ext", default=".txt", choices=[".txt", ".bin"])
    ap.add_argument("--max_d", type=int, default=192)
    ap.add_argument("--interval_scale", type=float, default=1.0)
    ap.add_argument("--theta0", type=float, default=1.0, help="min triangulation angle (deg)")
    ap.add_argument("--top_k", type=int, default=20, help="max neighbours kept per image")
    ap.add_argument("--min_shared", type=int, default=10, help="min shared tracks to keep a pair")
    ap.add_argument("--chunksize", type=int, default=512, help="accepted for compatibility with the reference CLI; unused")
    a = ap.parse_args(argv)
    os.makedirs(a.save_folder, exist_ok=True)
    convert(a.dense_folder, a.save_folder, a.model_ext, a.max_d, a.interval_scale, a.theta0, a.top_k, a.min_shared)


if __name__ == "__main__":
    main()
