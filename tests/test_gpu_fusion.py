"""GPU parity of the fusion stage: acmmp_fusion_* against the reference's SimpleFusionKernel
(ACMMP.cu:1664-1814, compiled unmodified into oracle/_ref/libacmmp_ref.so and fed through RunFusionCuda's texture set-up).
Inputs: per view a depth map (ground truth with noise and holes), world-frame normals, the grey image -- what
depths_geom.dmb / normals.dmb / images hold after the PatchMatch stages.  The kernel is deterministic: the same pixels must
produce a point (apart from pixels sitting on one of the three consistency thresholds), and the points must agree to 1e-4."""
import numpy as np
import pytest

import util
from util import dump

pytestmark = pytest.mark.gpu


def _inputs(scene, noise=0.002, holes=0.02, seed=9):
    from acmmp_b200 import synth
    rng = np.random.default_rng(seed)
    depths, normals = [], []
    for v in range(len(scene.images)):
        d = scene.depths_gt[v] * (1.0 + noise * rng.standard_normal(scene.depths_gt[v].shape)).astype(np.float32)
        d[rng.random(d.shape) < holes] = 0.0
        n = util.world_normals(scene, v, synth.gt_planes(scene, v))
        n += 0.02 * rng.standard_normal(n.shape).astype(np.float32)
        n /= np.linalg.norm(n, axis=-1, keepdims=True)
        depths.append(np.ascontiguousarray(d, np.float32))
        normals.append(np.ascontiguousarray(n, np.float32))
    return depths, normals


def _colour_images(scene):
    """B, G, R images with three DIFFERENT channels (a swapped pair would show): grey, shifted grey, inverted grey."""
    out = []
    for img in scene.images:
        g = np.clip(img, 0, 255).astype(np.uint8)
        out.append(np.ascontiguousarray(np.stack([g, np.roll(g, 5, axis=1), 255 - g], axis=-1)))
    return out


@pytest.mark.parametrize("colour", [False, True])
@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_fusion_matches_the_reference_kernel(model, colour):
    from acmmp_b200 import Fusion, synth
    from oracle.ref_driver import RefFusion
    scene = (synth.make_pinhole_scene(n_views=5, width=640, height=480, focal=500.0, seed=1) if model == "pinhole"
             else synth.make_sphere_scene(n_views=5, width=1024, height=512, seed=4))
    depths, normals = _inputs(scene)
    n = len(scene.images)
    mine = Fusion(n, 0)
    bgr = _colour_images(scene) if colour else None
    for v in range(n):
        mine.set_view(v, scene.cams[v], depths[v], normals[v], scene.images[v])
        if colour:
            mine.set_view_colour(v, bgr[v])
    ref = RefFusion(scene.cams, depths, normals, scene.images, colours=bgr)
    res = {}
    for r in range(n):
        src = list(scene.pairs[r][1])
        pa = mine.run(r, src)
        fa = mine.last_flags(r).astype(bool)
        pb, fb = ref.run(r, src)
        fb = fb.astype(bool)
        both = fa & fb
        # compacted arrays are in pixel order: rank of a pixel = number of flagged pixels before it
        ia = np.cumsum(fa.ravel()) - 1
        ib = np.cumsum(fb.ravel()) - 1
        sel = both.ravel()
        A, B = pa[ia[sel]], pb[ib[sel]]
        scale = np.abs(B[:, :3]).max()
        res[f"view{r}"] = dict(
            points_mine=int(fa.sum()), points_ref=int(fb.sum()), flags_equal=float((fa == fb).mean()),
            coord_within_1e4=float((np.abs(A[:, :3] - B[:, :3]) <= 1e-4 * scale).all(axis=1).mean()),
            normal_within_1e4=float((np.abs(A[:, 3:6] - B[:, 3:6]) <= 1e-4).all(axis=1).mean()),
            colour_within_half_level=float((np.abs(A[:, 6:] - B[:, 6:]) <= 0.5).all(axis=1).mean()),
            channels_differ=float((np.abs(B[:, 6] - B[:, 8]) > 1.0).mean()),
            kernel_ms_mine=mine.kernel_ms, kernel_ms_ref=ref.kernel_ms)
        assert len(pa) == int(fa.sum())
    dump(f"fusion_{model}" + ("_colour" if colour else ""), res)
    mine.close()
    ref.close()
    for k, v in res.items():
        assert v["points_ref"] > 1000, (k, v)
        assert v["flags_equal"] >= 0.9995, (k, v)
        assert v["coord_within_1e4"] >= 0.9999 and v["normal_within_1e4"] >= 0.9999, (k, v)
        assert v["colour_within_half_level"] >= 0.999, (k, v)
        if colour:
            assert v["channels_differ"] > 0.5, (k, v)          # the test images really have three different channels


def test_fusion_edge_cases():
    """A view without depth produces no point; missing source views (-1) are skipped; fewer than two consistent sources
    give nothing (the reference needs >= 3 views including the reference view, ACMMP.cu:1778); capacity too small reports
    the need."""
    from acmmp_b200 import Fusion, synth
    scene = synth.make_pinhole_scene(n_views=4, width=320, height=240, focal=250.0, seed=1)
    depths, normals = _inputs(scene, holes=0.0)
    f = Fusion(4, 0)
    for v in range(4):
        f.set_view(v, scene.cams[v], depths[v], normals[v], scene.images[v])
    full = f.run(0, [1, 2, 3])
    assert len(full) > 0.5 * 320 * 240
    assert len(f.run(0, [1])) == 0                       # one source: at most 2 consistent views
    assert len(f.run(0, [1, -1, -1])) == 0
    two = f.run(0, [1, 2])
    assert 0 < len(two) <= len(full)
    f.set_view(0, scene.cams[0], np.zeros_like(depths[0]), normals[0], scene.images[0])
    assert len(f.run(0, [1, 2, 3])) == 0
    f.close()


def test_ply_records_from_the_device_equal_the_host_writers(tmp_path):
    """acmmp_fusion_run_ply: the 27-byte vertex records packed on the device are byte for byte what the host writer
    (StoreColorPlyFileBinaryPointCloud, reference ACMMP.cpp:481-534) makes of the same points."""
    import ctypes as C
    from acmmp_b200 import Fusion, PKG_DIR, synth
    scene = synth.make_pinhole_scene(n_views=4, width=320, height=240, focal=250.0, seed=1)
    depths, normals = _inputs(scene, holes=0.01)
    bgr = _colour_images(scene)
    f = Fusion(4, 0)
    for v in range(4):
        f.set_view(v, scene.cams[v], depths[v], normals[v], scene.images[v])
        f.set_view_colour(v, bgr[v])
    src = list(scene.pairs[0][1])
    pts = f.run(0, src)
    rec = f.run_ply(0, src)
    f.close()
    assert len(pts) > 1000 and rec.shape == (len(pts), 27)
    host = C.CDLL(str(PKG_DIR / "lib" / "libacmmp_host.so"))
    path = tmp_path / "points.ply"
    assert host.acmmp_host_write_ply(str(path).encode(), np.ascontiguousarray(pts, np.float32).ctypes.data_as(C.POINTER(C.c_float)), len(pts)) == 0
    body = path.read_bytes().split(b"end_header\n", 1)[1]
    assert body == rec.tobytes()
