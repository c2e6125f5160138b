"""Golden fixtures for the COLMAP converter (acmmp_b200/colmap.py): a small synthetic COLMAP sparse model (text and
binary) per camera model + the cams/ and pair.txt the REFERENCE's own converter wrote for it.

Run in the build container only (it executes /root/reference/colmap2mvsnet_acm.py, a Python script, unmodified):
    python tests/golden/make_colmap_golden.py
The GPU box and the test suite read the committed files under tests/golden/colmap_{pinhole,sphere}/ only."""
import os
import shutil
import struct
import subprocess
import sys
import tempfile
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "acmmp-spherical_b200"))
REFERENCE = Path("/root/reference/colmap2mvsnet_acm.py")
ARGS = ["--top_k", "4", "--min_shared", "5", "--theta0", "1.0", "--max_d", "192"]
# a second set that takes the other branches: plane count from the one-pixel baseline (--max_d 0), pairs dropped by the
# triangulation-angle test and by the shared-track floor, an interval scale
ARGS_STRICT = ["--top_k", "3", "--min_shared", "60", "--theta0", "6.0", "--max_d", "0", "--interval_scale", "2"]


def rotmat_to_qvec(R):
    """(w, x, y, z) of a rotation matrix."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def build_model(scene, n_points, rng, sphere):
    """Random surface points of the scene's quads, observed by every view that sees them (no occlusion test: a sparse
    model only needs plausible tracks).  Returns (cameras, images, points) as plain lists."""
    from acmmp_b200 import MODEL_SPHERE
    H, W = scene.images[0].shape
    pts = []
    for _ in range(n_points):
        q = scene.quads[rng.integers(len(scene.quads))]
        pts.append(q.origin + q.u * rng.uniform(0, q.lu) + q.v * rng.uniform(0, q.lv))
    pts = np.asarray(pts)
    obs = [[] for _ in scene.images]                 # per image: (x, y, point id or -1)
    tracks = [[] for _ in pts]
    for v, (R, t) in enumerate(zip(scene.Rs, scene.ts)):
        Xc = pts @ R.T + t
        if sphere:
            d = np.linalg.norm(Xc, axis=1)
            x = np.arctan2(Xc[:, 0], Xc[:, 2]) / (2 * np.pi) * W + W / 2.0
            y = np.arcsin(Xc[:, 1] / d) / np.pi * H + H / 2.0
            ok = (x >= 0) & (x < W) & (y >= 0) & (y < H) & (rng.random(len(pts)) < 0.7)
        else:
            K = scene.Ks[v]
            x = K[0, 0] * Xc[:, 0] / Xc[:, 2] + K[0, 2]
            y = K[1, 1] * Xc[:, 1] / Xc[:, 2] + K[1, 2]
            ok = (Xc[:, 2] > 0) & (x >= 0) & (x < W) & (y >= 0) & (y < H) & (rng.random(len(pts)) < 0.8)
        for p in np.nonzero(ok)[0]:
            tracks[p].append((v + 1, len(obs[v])))
            obs[v].append((float(x[p]), float(y[p]), int(p) + 1))
        for _ in range(5):                                   # a few observations without a 3-D point
            obs[v].append((float(rng.uniform(0, W)), float(rng.uniform(0, H)), -1))
    if sphere:
        cams = [(1, "SPHERE", W, H, [1.0, W / 2.0, H / 2.0])]
    else:
        K = scene.Ks[0]
        cams = [(1, "PINHOLE", W, H, [K[0, 0], K[1, 1], K[0, 2], K[1, 2]])]
    # COLMAP image ids deliberately not contiguous: the converter re-indexes by sorted id
    images = [(3 * v + 2, rotmat_to_qvec(R), t, 1, "view_%02d.%s" % (v, "jpg" if v % 2 else "png"), obs[v]) for v, (R, t) in enumerate(zip(scene.Rs, scene.ts))]
    # tracks refer to image ids
    points = [(p + 1, pts[p], [(3 * (iid - 1) + 2, k) for iid, k in tracks[p]]) for p in range(len(pts)) if len(tracks[p]) >= 2]
    keep = {p[0] for p in points}
    images = [(iid, q, t, cid, name, [(x, y, (pid if pid in keep else -1)) for x, y, pid in o]) for iid, q, t, cid, name, o in images]
    return cams, images, points


def write_text(folder, cams, images, points):
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "cameras.txt"), "w") as f:
        f.write("# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n")
        for cid, model, w, h, params in cams:
            f.write(f"{cid} {model} {w} {h} " + " ".join(repr(float(p)) for p in params) + "\n")
    with open(os.path.join(folder, "images.txt"), "w") as f:
        f.write("# Image list with two lines of data per image:\n")
        for iid, q, t, cid, name, o in images:
            f.write(f"{iid} " + " ".join(repr(float(v)) for v in q) + " " + " ".join(repr(float(v)) for v in t) + f" {cid} {name}\n")
            f.write(" ".join(f"{x!r} {y!r} {pid}" for x, y, pid in o) + "\n")
    with open(os.path.join(folder, "points3D.txt"), "w") as f:
        f.write("# 3D point list with one line of data per point:\n")
        for pid, xyz, track in points:
            f.write(f"{pid} " + " ".join(repr(float(v)) for v in xyz) + " 128 128 128 0.5 " + " ".join(f"{i} {k}" for i, k in track) + "\n")


def write_binary(folder, cams, images, points):
    os.makedirs(folder, exist_ok=True)
    ids = {"PINHOLE": 1, "SPHERE": 11}
    with open(os.path.join(folder, "cameras.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(cams)))
        for cid, model, w, h, params in cams:
            f.write(struct.pack("<iiQQ", cid, ids[model], w, h) + struct.pack("<%dd" % len(params), *params))
    with open(os.path.join(folder, "images.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(images)))
        for iid, q, t, cid, name, o in images:
            f.write(struct.pack("<idddddddi", iid, *q, *t, cid) + name.encode() + b"\x00" + struct.pack("<Q", len(o)))
            for x, y, pid in o:
                f.write(struct.pack("<ddq", x, y, pid))
    with open(os.path.join(folder, "points3D.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(points)))
        for pid, xyz, track in points:
            f.write(struct.pack("<QdddBBBd", pid, *xyz, 128, 128, 128, 0.5) + struct.pack("<Q", len(track)))
            for i, k in track:
                f.write(struct.pack("<ii", i, k))


def main():
    from acmmp_b200 import synth
    assert REFERENCE.exists(), "the reference converter is only available in the build container"
    rng = np.random.default_rng(11)
    for name in ("pinhole", "sphere"):
        sphere = name == "sphere"
        scene = (synth.make_sphere_scene(n_views=8, width=256, height=128, seed=4) if sphere
                 else synth.make_pinhole_scene(n_views=8, width=160, height=120, focal=130.0, seed=1))
        cams, images, points = build_model(scene, 600, rng, sphere)
        out = ROOT / "tests" / "golden" / f"colmap_{name}"
        shutil.rmtree(out, ignore_errors=True)
        dense = out / "dense"
        write_text(dense / "sparse", cams, images, points)
        write_binary(dense / "sparse", cams, images, points)
        os.makedirs(dense / "images", exist_ok=True)
        for (iid, q, t, cid, fname, o), img in zip(images, scene.images):
            small = cv2.resize(img.astype(np.uint8), (32, 24))                 # the converter only copies / re-encodes them
            cv2.imwrite(str(dense / "images" / fname), small)
        for ext, args, sub in ((".txt", ARGS, "expected"), (".bin", ARGS, "expected_bin"), (".txt", ARGS_STRICT, "expected_strict")):
            with tempfile.TemporaryDirectory() as tmp:
                r = subprocess.run([sys.executable, str(REFERENCE), "--dense_folder", str(dense), "--save_folder", tmp, "--model_ext", ext] + args,
                                   capture_output=True, text=True)
                assert r.returncode == 0, r.stderr[-2000:]
                exp = out / sub
                shutil.copytree(os.path.join(tmp, "cams"), exp / "cams")
                shutil.copyfile(os.path.join(tmp, "pair.txt"), exp / "pair.txt")
                (exp / "images.lst").write_text("\n".join(sorted(os.listdir(os.path.join(tmp, "images")))) + "\n")
        print(name, "views", len(images), "points", len(points), "->", out)


if __name__ == "__main__":
    main()
