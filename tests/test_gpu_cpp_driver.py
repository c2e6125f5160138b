"""GPU test of the C++ host side: the `acmmp_b200` driver (reference: ./ACMMP dense_folder, main.cpp:392-482) on a
synthetic dense folder, through the reference's own on-disk contract."""
import json
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

import util

ROOT = Path(__file__).resolve().parent.parent
DRIVER = ROOT / "acmmp-spherical_b200" / "lib" / "acmmp_b200"


def _read_dmb(path):
    raw = open(path, "rb").read()
    t, h, w, nb = struct.unpack("<4i", raw[:16])
    assert t == 1
    a = np.frombuffer(raw[16:], np.float32)
    return a.reshape(h, w) if nb == 1 else a.reshape(h, w, nb)


@pytest.mark.gpu
def test_cpp_driver_runs_the_reference_schedule_on_a_dense_folder(tmp_path):
    from acmmp_b200 import synth
    assert DRIVER.exists(), "build the host side first (__graft_entry__.build())"
    # 1100 px -> two pyramid levels (550 and 1100): JBU + hierarchy + prior + 2 geometric rounds per level
    scene = synth.make_pinhole_scene(n_views=4, width=1100, height=820, focal=950.0, seed=3)
    synth.write_dense_folder(scene, str(tmp_path), pgm=True)
    # colour twins for the fused points (the reference reads the .jpg in colour, ACMMP.cu:1862): B = g, G = shifted g, R = 255 - g
    import cv2
    for v, img in enumerate(scene.images):
        g = np.clip(img, 0, 255).astype(np.uint8)
        assert cv2.imwrite(str(tmp_path / "images" / ("%08d.ppm" % v)), np.stack([g, np.roll(g, 5, axis=1), 255 - g], axis=-1))
    r = subprocess.run([str(DRIVER), str(tmp_path), "--seed", "7"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    assert summary["views"] == 4 and summary["kernel_ms"] > 0
    res = {}
    for v in range(4):
        folder = tmp_path / "ACMMP" / ("2333_%08d" % v)
        for name in ("depths.dmb", "depths_geom.dmb", "normals.dmb", "costs.dmb"):
            assert (folder / name).exists(), name
        depth = _read_dmb(folder / "depths_geom.dmb")
        normals = _read_dmb(folder / "normals.dmb")
        costs = _read_dmb(folder / "costs.dmb")
        gt = scene.depths_gt[v]
        assert depth.shape == gt.shape and normals.shape == gt.shape + (3,) and costs.shape == gt.shape
        ok = np.abs(depth - gt) / gt <= 0.01
        res[f"view{v}_within_1pct_of_gt"] = float(ok[8:-8, 8:-8].mean())
        res[f"view{v}_unit_normals"] = float((np.abs(np.linalg.norm(normals, axis=-1) - 1) < 1e-3).mean())
    # fusion (RunFusionCuda, main.cpp:478-479): ACMMP/ACMM_model_cuda_5.ply, 27 bytes per point after the header
    ply = (tmp_path / "ACMMP" / "ACMM_model_cuda_5.ply").read_bytes()
    head, body = ply.split(b"end_header\n", 1)
    n_points = int([l for l in head.decode().splitlines() if l.startswith("element vertex")][0].split()[-1])
    assert n_points == summary["fusion_points"] and len(body) == 27 * n_points
    pts = np.frombuffer(body, np.dtype([("xyz", "<f4", 3), ("n", "<f4", 3), ("rgb", "u1", 3)]))
    res["fusion_points"] = n_points
    res["fusion_points_per_pixel"] = n_points / (4.0 * scene.depths_gt[0].size)
    res["fusion_unit_normals"] = float((np.abs(np.linalg.norm(pts["n"], axis=-1) - 1) < 1e-3).mean())
    # PLY order red, green, blue: red = mean of (255 - g), blue = mean of g over the consistent views => red + blue = 255
    rb = pts["rgb"][:, 0].astype(int) + pts["rgb"][:, 2].astype(int)
    res["fusion_red_plus_blue_is_255"] = float((np.abs(rb - 255) <= 2).mean())
    res["fusion_colour_spread"] = float(np.abs(pts["rgb"][:, 0].astype(int) - pts["rgb"][:, 2].astype(int)).mean())
    # every fused point projects into view 0 at a depth that agrees with that view's ground-truth depth
    R, t, K = scene.Rs[0], scene.ts[0], scene.Ks[0]
    Xc = pts["xyz"].astype(np.float64) @ R.T + t
    uv = (Xc @ K.T)[:, :2] / Xc[:, 2:3]
    ui, vi = np.rint(uv[:, 0]).astype(int), np.rint(uv[:, 1]).astype(int)
    gt0 = scene.depths_gt[0]
    inside = (ui >= 8) & (ui < gt0.shape[1] - 8) & (vi >= 8) & (vi < gt0.shape[0] - 8) & (Xc[:, 2] > 0)
    rel = np.abs(Xc[inside, 2] - gt0[vi[inside], ui[inside]]) / gt0[vi[inside], ui[inside]]
    res["fusion_points_within_2pct_of_gt_surface"] = float((rel <= 0.02).mean())
    res.update(summary)
    util.dump("cpp_driver", res)
    for v in range(4):
        assert res[f"view{v}_within_1pct_of_gt"] > 0.85, res
        assert res[f"view{v}_unit_normals"] > 0.99, res
    assert res["fusion_points_per_pixel"] > 0.3 and res["fusion_unit_normals"] > 0.99, res
    assert res["fusion_red_plus_blue_is_255"] > 0.99 and res["fusion_colour_spread"] > 10, res      # colours, in the right channels
    assert res["fusion_points_within_2pct_of_gt_surface"] > 0.8, res      # occluded points excepted


@pytest.mark.gpu
def test_cpp_driver_resident_schedule_and_gpu_prior_write_the_same_maps_as_the_file_chained_one(tmp_path):
    """`--resident 1` (stage state, JBU hand-over and neighbour depth maps stay on the device, images
    are read once per level) and `--gpu-prior 1` (N2: support points, plane fit, rasteriser and depth-range test of the
    planar-prior stage on the device, written with round-to-nearest intrinsics in the host code's evaluation order)
    run the same arithmetic on the same inputs in the same order as the reference's file-chained schedule with its CPU
    prior stage, so the final maps must be IDENTICAL bit for bit -- a size-independent property, no tolerance."""
    import shutil
    from acmmp_b200 import synth
    assert DRIVER.exists(), "build the host side first (__graft_entry__.build())"
    scene = synth.make_pinhole_scene(n_views=4, width=1100, height=820, focal=950.0, seed=5)
    base = tmp_path / "files"
    base.mkdir()
    synth.write_dense_folder(scene, str(base), pgm=True)
    variants = {"files": ("0", "0"), "resident": ("1", "0"), "files_gpu_prior": ("0", "1"), "resident_gpu_prior": ("1", "1")}
    folders, out = {}, {}
    for name in variants:
        folders[name] = base if name == "files" else tmp_path / name
        if name != "files":
            shutil.copytree(base, folders[name])
    for name, (resident, gpu_prior) in variants.items():
        r = subprocess.run([str(DRIVER), str(folders[name]), "--seed", "11", "--resident", resident, "--gpu-prior", gpu_prior],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
        assert out[name]["mode"] == ("resident" if resident == "1" else "files") and out[name]["gpu_prior"] == int(gpu_prior), out[name]
    res = dict(out)
    for name in variants:
        if name == "files":
            continue
        for v in range(4):
            for dmb in ("depths.dmb", "depths_geom.dmb", "normals.dmb", "costs.dmb"):
                x = _read_dmb(folders["files"] / "ACMMP" / ("2333_%08d" % v) / dmb)
                y = _read_dmb(folders[name] / "ACMMP" / ("2333_%08d" % v) / dmb)
                assert x.shape == y.shape, (name, v, dmb)
                # bit patterns: costs hold NaN where no view was selected (0 / 0, as in the reference), and NaN != NaN
                res[f"{name}_view{v}_{dmb}_identical"] = bool(np.array_equal(x.view(np.uint32), y.view(np.uint32)))
        # the fused point cloud: the resident schedules feed the fusion from the maps still on the device, the file-chained
        # ones from the .dmb files -- the same maps, so the same bytes
        res[f"{name}_ply_identical"] = bool((folders["files"] / "ACMMP" / "ACMM_model_cuda_5.ply").read_bytes()
                                            == (folders[name] / "ACMMP" / "ACMM_model_cuda_5.ply").read_bytes())
    util.dump("cpp_driver_resident", res)
    bad = [k for k, v in res.items() if k.endswith("_identical") and not v]
    assert not bad, bad
    # wall times and CPU-side prior times of the four schedules are in the metrics dump (process start-up and the page
    # cache dominate the wall time at this size: not asserted)


@pytest.mark.gpu
def test_cpp_driver_on_an_equirectangular_dense_folder(tmp_path):
    """The fork's SPHERE camera model through the on-disk contract (`intrinsic / SPHERE / f cx cy` cam files,
    reference ACMMP.cpp:172-192; BASELINE config 4 at a small size, two pyramid levels): the file-chained and the
    GPU-resident schedule write bit-identical maps; the device prior stage (its sin / cos differ from the host's in the
    last bit) agrees with them within the full-map tolerance; depths match the ground truth."""
    import shutil
    from acmmp_b200 import synth
    assert DRIVER.exists(), "build the host side first (__graft_entry__.build())"
    scene = synth.make_sphere_scene(n_views=4, width=1200, height=600, seed=4)
    base = tmp_path / "files"
    base.mkdir()
    synth.write_dense_folder(scene, str(base), pgm=True)
    variants = {"files": ("0", "0"), "resident": ("1", "0"), "resident_gpu_prior": ("1", "1")}
    folders, res = {}, {}
    for name, (resident, gpu_prior) in variants.items():
        folders[name] = base if name == "files" else tmp_path / name
        if name != "files":
            shutil.copytree(base, folders[name])
        r = subprocess.run([str(DRIVER), str(folders[name]), "--seed", "3", "--resident", resident, "--gpu-prior", gpu_prior],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res[name] = json.loads(r.stdout.strip().splitlines()[-1])
    for v in range(4):
        ref = {d: _read_dmb(folders["files"] / "ACMMP" / ("2333_%08d" % v) / d) for d in ("depths_geom.dmb", "normals.dmb", "costs.dmb")}
        for d, x in ref.items():
            y = _read_dmb(folders["resident"] / "ACMMP" / ("2333_%08d" % v) / d)
            res[f"resident_view{v}_{d}_identical"] = bool(x.shape == y.shape and np.array_equal(x.view(np.uint32), y.view(np.uint32)))
        z = _read_dmb(folders["resident_gpu_prior"] / "ACMMP" / ("2333_%08d" % v) / "depths_geom.dmb")
        x = ref["depths_geom.dmb"]
        res[f"gpu_prior_view{v}_within_1pct_of_files"] = float((np.abs(z - x) <= 0.01 * np.abs(x)).mean())
        gt = scene.depths_gt[v]
        assert x.shape == gt.shape
        res[f"view{v}_within_1pct_of_gt"] = float((np.abs(x - gt) / gt <= 0.01)[8:-8, 8:-8].mean())
    util.dump("cpp_driver_sphere", res)
    for v in range(4):
        for d in ("depths_geom.dmb", "normals.dmb", "costs.dmb"):
            assert res[f"resident_view{v}_{d}_identical"], res
        assert res[f"gpu_prior_view{v}_within_1pct_of_files"] > 0.99, res
        assert res[f"view{v}_within_1pct_of_gt"] > 0.85, res


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("levels", [1, 2])
def test_cpp_driver_multi_gpu_shards_views_and_exchanges_depth_maps(tmp_path, levels):
    """`--gpus 2`: the reference views dealt round-robin to two devices, one host thread per device, the
    neighbours' depth maps pulled over NVLink at the two exchange points of a level.  Photometric and prior stages are
    per-view work: on a one-level scene a sharded run holds the SAME maps after them as the one-device run (depths.dmb
    bit-identical).  The geometric rounds read other views' maps: round 1 is Gauss-Seidel inside a device and Jacobi across
    devices, so the final maps (and everything a finer level builds on them) agree within the full-map tolerance, not bit
    for bit.  Needs two GPUs (skipped otherwise)."""
    import shutil
    from acmmp_b200 import synth
    if _device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    w, h, f = (900, 680, 780.0) if levels == 1 else (1100, 820, 950.0)
    scene = synth.make_pinhole_scene(n_views=5, width=w, height=h, focal=f, seed=5)
    one = tmp_path / "one"
    one.mkdir()
    synth.write_dense_folder(scene, str(one), pgm=True)
    two = tmp_path / "two"
    shutil.copytree(one, two)
    out = {}
    for name, folder, gpus in (("one", one, "1"), ("two", two, "2")):
        r = subprocess.run([str(DRIVER), str(folder), "--seed", "11", "--resident", "1", "--gpu-prior", "1", "--gpus", gpus],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["one"]["gpus"] == 1 and out["two"]["gpus"] == 2, out
    res = dict(out)
    for v in range(5):
        a = {d: _read_dmb(one / "ACMMP" / ("2333_%08d" % v) / d) for d in ("depths.dmb", "depths_geom.dmb", "normals.dmb")}
        b = {d: _read_dmb(two / "ACMMP" / ("2333_%08d" % v) / d) for d in ("depths.dmb", "depths_geom.dmb", "normals.dmb")}
        res[f"view{v}_prior_stage_depths_identical"] = bool(np.array_equal(a["depths.dmb"].view(np.uint32), b["depths.dmb"].view(np.uint32)))
        rel = np.abs(a["depths_geom.dmb"] - b["depths_geom.dmb"]) / np.maximum(np.abs(a["depths_geom.dmb"]), 1e-9)
        res[f"view{v}_final_depth_within_1pct"] = float((rel <= 0.01)[8:-8, 8:-8].mean())
        gt = scene.depths_gt[v]
        res[f"view{v}_within_1pct_of_gt"] = float((np.abs(b["depths_geom.dmb"] - gt) / gt <= 0.01)[8:-8, 8:-8].mean())
    # the two-device run fed its fusion from the maps on the devices (peer copies to the first one); fusing its .dmb files
    # again (--fusion-only) must give the same bytes
    ply = two / "ACMMP" / "ACMM_model_cuda_5.ply"
    resident_ply = ply.read_bytes()
    r = subprocess.run([str(DRIVER), str(two), "--fusion-only", "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res["ply_of_resident_maps_equals_ply_of_files"] = bool(ply.read_bytes() == resident_ply)
    res["fusion_points_two"] = out["two"]["fusion_points"]
    util.dump(f"cpp_driver_multi_gpu_{levels}_levels", res)
    assert res["ply_of_resident_maps_equals_ply_of_files"] and res["fusion_points_two"] > 1000, res
    for v in range(5):
        if levels == 1:
            assert res[f"view{v}_prior_stage_depths_identical"], res
        assert res[f"view{v}_final_depth_within_1pct"] > 0.97, res
        assert res[f"view{v}_within_1pct_of_gt"] > 0.85, res


@pytest.mark.gpu
@pytest.mark.parametrize("colour", [False, True])
def test_jpeg_views_are_decoded_on_the_device_like_imread_grayscale(tmp_path, colour):
    """The reference reads images/%08d.jpg with cv::imread(IMREAD_GRAYSCALE) (ACMMP.cpp:578); the host side decodes the luma
    plane with nvJPEG.  Both are the JPEG's Y channel: equal up to the IDCT's rounding (libjpeg vs nvJPEG: a grey level or
    two on a few pixels).  Grey and colour files, odd sizes (chroma subsampling with partial blocks)."""
    import ctypes as C
    import cv2
    from acmmp_b200 import synth
    lib = C.CDLL(str(ROOT / "acmmp-spherical_b200" / "lib" / "libacmmp_host.so"))
    scene = synth.make_pinhole_scene(n_views=2, width=397, height=251, focal=300.0, seed=8)
    (tmp_path / "images").mkdir()
    res = {}
    for i, img in enumerate(scene.images):
        g = img.astype(np.uint8)
        src = np.stack([g, np.roll(g, 3, axis=1), 255 - g], axis=-1) if colour else g
        path = str(tmp_path / "images" / ("%08d.jpg" % i))
        assert cv2.imwrite(path, src, [cv2.IMWRITE_JPEG_QUALITY, 92])
        want = cv2.imread(path, cv2.IMREAD_GRAYSCALE).astype(np.float32)
        got = np.zeros(want.shape, np.float32)
        w, h = C.c_int(0), C.c_int(0)
        rc = lib.acmmp_host_load_grey(str(tmp_path).encode(), C.c_int(i), got.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(got.size),
                                      C.byref(w), C.byref(h))
        assert rc == 0 and (h.value, w.value) == want.shape
        d = np.abs(got - want)
        res[f"view{i}"] = dict(max_abs=float(d.max()), mean_abs=float(d.mean()), equal=float((d == 0).mean()))
        assert d.max() <= 3 and d.mean() < 0.35
    util.dump("jpeg_decode_" + ("colour" if colour else "grey"), res)
