"""CPU tests of the C++ host side (acmmp-spherical_b200/host): on-disk contract, image resampling,
Delaunay stand-in.  No GPU work: the helpers are reached through the C hooks of libacmmp_host.so."""
import ctypes as C
import struct
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST_LIB = ROOT / "acmmp-spherical_b200" / "lib" / "libacmmp_host.so"


@pytest.fixture(scope="module")
def host():
    if not HOST_LIB.exists():
        pytest.fail(f"{HOST_LIB} is missing: run __graft_entry__.build()")
    return C.CDLL(str(HOST_LIB))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def test_dmb_round_trip_and_layout(host, tmp_path):
    """.dmb = int32 type(1), h, w, nb + float32 payload, row major (reference ACMMP.cpp:352-479)."""
    rng = np.random.default_rng(0)
    d = rng.uniform(1, 9, (7, 11)).astype(np.float32)
    n = rng.standard_normal((7, 11, 3)).astype(np.float32)
    pd, pn = str(tmp_path / "depths.dmb").encode(), str(tmp_path / "normals.dmb").encode()
    assert host.acmmp_host_write_depth_dmb(pd, _fp(d), 11, 7) == 0
    assert host.acmmp_host_write_normal_dmb(pn, _fp(n), 11, 7) == 0
    raw = open(pd, "rb").read()
    assert struct.unpack("<4i", raw[:16]) == (1, 7, 11, 1)
    assert np.array_equal(np.frombuffer(raw[16:], np.float32).reshape(7, 11), d)
    raw = open(pn, "rb").read()
    assert struct.unpack("<4i", raw[:16]) == (1, 7, 11, 3)
    assert np.array_equal(np.frombuffer(raw[16:], np.float32).reshape(7, 11, 3), n)
    out = np.zeros((7, 11), np.float32)
    w, h = C.c_int(), C.c_int()
    assert host.acmmp_host_read_depth_dmb(pd, _fp(out), out.size, C.byref(w), C.byref(h)) == 0
    assert (w.value, h.value) == (11, 7) and np.array_equal(out, d)
    outn = np.zeros((7, 11, 3), np.float32)
    assert host.acmmp_host_read_normal_dmb(pn, _fp(outn), outn.size, C.byref(w), C.byref(h)) == 0
    assert np.array_equal(outn, n)
    # a depth file is not a normal file
    assert host.acmmp_host_read_normal_dmb(pd, _fp(outn), outn.size, C.byref(w), C.byref(h)) != 0
    assert host.acmmp_host_read_depth_dmb(str(tmp_path / "missing.dmb").encode(), _fp(out), out.size, C.byref(w), C.byref(h)) != 0


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_camera_files_written_by_the_generator_parse_back(host, tmp_path, model):
    """cams/%08d_cam.txt as the reference reads them (ACMMP.cpp:146-209), including the PINHOLE depth-line quirk."""
    import acmmp_b200
    from acmmp_b200 import synth
    scene = (synth.make_pinhole_scene(n_views=3, width=64, height=48, focal=60.0, seed=1) if model == "pinhole"
             else synth.make_sphere_scene(n_views=3, width=64, height=32, seed=4))
    synth.write_dense_folder(scene, str(tmp_path), pgm=True)
    for i, ref in enumerate(scene.cams):
        cam = acmmp_b200.Camera()
        assert host.acmmp_host_read_camera(str(tmp_path / "cams" / ("%08d_cam.txt" % i)).encode(), C.byref(cam)) == 0
        assert cam.model == ref.model
        np.testing.assert_allclose(list(cam.R), list(ref.R), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(list(cam.t), list(ref.t), rtol=1e-6, atol=1e-7)
        if model == "pinhole":
            np.testing.assert_allclose(list(cam.K), list(ref.K), rtol=1e-6)
        else:
            np.testing.assert_allclose(list(cam.params)[:3], list(ref.params)[:3], rtol=1e-6)
        assert abs(cam.depth_min - ref.depth_min) <= 1e-5 * ref.depth_min
        assert abs(cam.depth_max - ref.depth_max) <= 1e-5 * ref.depth_max
    # pair.txt: every view listed with its kept sources
    assert host.acmmp_host_pair_count(str(tmp_path).encode()) == sum(1000 + len(s) for _, s in scene.pairs)
    # the lossless twin is what the loader returns
    img = np.zeros(scene.images[0].shape, np.float32)
    w, h = C.c_int(), C.c_int()
    assert host.acmmp_host_load_grey(str(tmp_path).encode(), 0, _fp(img), img.size, C.byref(w), C.byref(h)) == 0
    assert np.array_equal(img, scene.images[0].astype(np.uint8).astype(np.float32))


@pytest.mark.parametrize("shape,new", [((96, 128), (48, 64)), ((97, 131), (41, 77)), ((50, 60), (50, 60)), ((64, 64), (23, 57))])
def test_resize_matches_cv_inter_linear(host, shape, new):
    """InuputInitialization resamples with cv::resize(..., INTER_LINEAR) (ACMMP.cpp:624); the C++ host restates it."""
    import cv2
    rng = np.random.default_rng(3)
    src = rng.uniform(0, 255, shape).astype(np.float32)
    dst = np.zeros(new, np.float32)
    assert host.acmmp_host_resize_linear(_fp(src), shape[1], shape[0], _fp(dst), new[1], new[0]) == 0
    ref = cv2.resize(src, (new[1], new[0]), interpolation=cv2.INTER_LINEAR)
    assert np.abs(dst - ref).max() <= 2e-4 * 255


def _circumcircle_empty(pts, tris):
    """Delaunay property, exact in integers: no point strictly inside any triangle's circumcircle."""
    P = pts.astype(object)
    bad = 0
    for a, b, c in tris[:: max(1, len(tris) // 150)]:
        ax, ay = P[a]; bx, by = P[b]; cx, cy = P[c]
        if (bx - ax) * (cy - ay) - (by - ay) * (cx - ax) < 0:
            bx, by, cx, cy = cx, cy, bx, by
        for d in range(len(P)):
            if d in (a, b, c):
                continue
            dx, dy = P[d]
            m = [[ax - dx, ay - dy], [bx - dx, by - dy], [cx - dx, cy - dy]]
            r = [u * u + v * v for u, v in m]
            det = (m[0][0] * (m[1][1] * r[2] - r[1] * m[2][1]) - m[0][1] * (m[1][0] * r[2] - r[1] * m[2][0])
                   + r[0] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]))
            bad += det > 0
    return bad


def test_delaunay_stand_in_is_a_delaunay_triangulation(host):
    """DelaunayIndices replaces cv::Subdiv2D (ACMMP.cpp:932-954): same triangle count as scipy's Qhull on points in
    general position and the empty-circumcircle property on near-grid support points (co-circular ties included)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(5)
    # (1) support-point-like input: one point per 5x5 cell, scan order col-major like GetSupportPoints
    cells = [(cx, cy) for cx in range(40) for cy in range(30)]
    pts = np.array([(5 * cx + rng.integers(0, 5), 5 * cy + rng.integers(0, 5)) for cx, cy in cells if rng.random() < 0.8], np.int32)
    out = np.zeros((4 * len(pts), 3), np.int32)
    nt = host.acmmp_host_delaunay(pts.ctypes.data_as(C.POINTER(C.c_int32)), len(pts), out.ctypes.data_as(C.POINTER(C.c_int32)), len(out))
    tris = out[:nt]
    assert nt > len(pts) and tris.min() >= 0 and tris.max() < len(pts)
    assert _circumcircle_empty(pts, tris) == 0
    # Qhull triangulates the convex hull completely; the big enclosing triangle may leave out a few thin hull triangles
    ref = Delaunay(pts.astype(np.float64)).simplices
    assert abs(nt - len(ref)) <= 0.02 * len(ref) + 8
    def total_area(t):
        u, v = (pts[t[:, 1]] - pts[t[:, 0]]).astype(np.float64), (pts[t[:, 2]] - pts[t[:, 0]]).astype(np.float64)
        return 0.5 * np.abs(u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]).sum()
    area, area_ref = total_area(tris), total_area(ref)
    assert 0.99 * area_ref <= area <= area_ref * (1 + 1e-9)
    # (2) exact grid (every quadruple co-circular) and duplicates must not break it
    gx, gy = np.meshgrid(np.arange(0, 60, 5), np.arange(0, 40, 5))
    grid = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.int32)
    grid = np.concatenate([grid, grid[:7]])
    out = np.zeros((4 * len(grid), 3), np.int32)
    nt = host.acmmp_host_delaunay(grid.ctypes.data_as(C.POINTER(C.c_int32)), len(grid), out.ctypes.data_as(C.POINTER(C.c_int32)), len(out))
    assert nt == 2 * 11 * 7
    assert _circumcircle_empty(grid, out[:nt]) == 0


def test_delaunay_scales_to_full_resolution_point_counts(host):
    """~270 k support points at 3200x2130: the walk-located insertion must stay near-linear."""
    import time
    rng = np.random.default_rng(9)
    cx, cy = np.meshgrid(np.arange(640), np.arange(426), indexing="ij")
    pts = np.stack([5 * cx.ravel() + rng.integers(0, 5, cx.size), 5 * cy.ravel() + rng.integers(0, 5, cx.size)], 1).astype(np.int32)
    out = np.zeros((2 * len(pts) + 16, 3), np.int32)
    t0 = time.perf_counter()
    nt = host.acmmp_host_delaunay(pts.ctypes.data_as(C.POINTER(C.c_int32)), len(pts), out.ctypes.data_as(C.POINTER(C.c_int32)), len(out))
    dt = time.perf_counter() - t0
    assert 1.9 * len(pts) < nt < 2.0 * len(pts)
    assert dt < 20.0, f"{dt:.1f} s for {len(pts)} points"


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_cpu_planar_prior_stage_agrees_with_the_numpy_twin(host, model):
    """PlanarPriorCpu (host/acmmp_host.cpp: main.cpp:113-185 with the reference's stepping rasteriser and the exact-integer
    Delaunay stand-in) against acmmp_b200/prior.py (cv2.Subdiv2D + cv2.fillConvexPoly + closed-form planes): the Delaunay
    triangulation of generic points is unique, so away from triangle edges every pixel must get the same prior plane."""
    import sys
    sys.path.insert(0, str(ROOT / "acmmp-spherical_b200"))
    from acmmp_b200 import synth
    from acmmp_b200.prior import planar_prior
    rng = np.random.default_rng(3)
    if model == "pinhole":
        sc = synth.make_pinhole_scene(n_views=2, width=200, height=150, focal=160.0, seed=5)
    else:
        sc = synth.make_sphere_scene(n_views=2, width=256, height=128, seed=4)
    cam = sc.cams[0]
    gt = sc.depths_gt[0].astype(np.float32)
    H, W = gt.shape
    depths = np.ascontiguousarray(gt * (1.0 + 0.002 * rng.standard_normal((H, W)))).astype(np.float32)
    costs = np.ascontiguousarray(rng.uniform(0.0, 0.09, (H, W))).astype(np.float32)      # every cell yields a generic support point
    dmin, dmax = float(gt.min() * 0.6), float(gt.max() * 1.2)
    masks = np.zeros((H, W), np.float32)
    cap = 2 * ((W + 4) // 5) * ((H + 4) // 5) + 16
    params = np.zeros((cap, 4), np.float32)
    n = host.acmmp_host_planar_prior(C.byref(cam), W, H, _fp(depths), _fp(costs), C.c_float(dmin), C.c_float(dmax), _fp(masks), _fp(params), cap)
    assert 0 < n <= cap and masks.max() == n
    p_ref, m_ref = planar_prior(cam, depths, costs, dmin, dmax)
    assert abs(len(p_ref) - n) <= 0.02 * n                       # co-circular quadruples may be split either way
    assert abs((masks > 0).mean() - (m_ref > 0).mean()) < 0.03
    # the same triangles => the same set of planes (ids differ: the two triangulators list triangles in different orders)
    from scipy.spatial import cKDTree
    scale = np.abs(p_ref).max(axis=0)
    d_mine, _ = cKDTree(params[:n] / scale).query(p_ref / scale)
    d_ref, _ = cKDTree(p_ref / scale).query(params[:n] / scale)
    assert (d_mine < 1e-4).mean() > 0.98 and (d_ref < 1e-4).mean() > 0.98, ((d_mine < 1e-4).mean(), (d_ref < 1e-4).mean())
    # per pixel: interior pixels carry the same plane; pixels on triangle edges (a large share for 5-pixel triangles) go to
    # either neighbour depending on the rasteriser (reference stepping loop vs cv2 scan conversion), whose planes are close
    both = (masks > 0) & (m_ref > 0)
    assert both.mean() > 0.85
    mine = params[masks[both].astype(np.int64) - 1]
    theirs = p_ref[m_ref[both].astype(np.int64) - 1]
    same = np.all(np.abs(mine - theirs) <= 1e-4 + 1e-4 * np.abs(theirs), axis=1)
    assert same.mean() > 0.55, same.mean()


def test_ply_writer_writes_the_references_binary_layout(tmp_path):
    """StoreColorPlyFileBinaryPointCloud (reference ACMMP.cpp:481-534): header lines, 27-byte records (6 float + 3 uchar),
    colour stored as (b, g, r) in PointList and written red-green-blue, non-finite coordinates zeroed."""
    import ctypes as C
    import struct
    lib = C.CDLL(str(ROOT / "acmmp-spherical_b200" / "lib" / "libacmmp_host.so"))
    pts = np.array([[1.0, 2.0, 3.0, 0.0, 0.0, 1.0, 10.7, 20.2, 30.9],
                    [np.inf, 5.0, 6.0, 0.6, 0.0, 0.8, 255.0, 128.0, 0.0]], np.float32)
    path = tmp_path / "cloud.ply"
    assert lib.acmmp_host_write_ply(str(path).encode(), pts.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(len(pts))) == 0
    raw = path.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 2"]
    assert lines[3:] == ["property float x", "property float y", "property float z", "property float nx", "property float ny",
                         "property float nz", "property uchar red", "property uchar green", "property uchar blue"]
    assert len(body) == 2 * 27
    a = struct.unpack("<6f3B", body[:27])
    b = struct.unpack("<6f3B", body[27:])
    assert a == (1.0, 2.0, 3.0, 0.0, 0.0, 1.0, 30, 20, 10)                    # truncation, r = color.z, b = color.x
    assert b[:3] == (0.0, 0.0, 0.0) and b[6:] == (0, 128, 255)
    assert abs(b[3] - 0.6) < 1e-6 and abs(b[5] - 0.8) < 1e-6


@pytest.mark.parametrize("kind", ["grid", "ragged", "clustered"])
def test_delaunay_triangle_set_equals_cv2_subdiv2d_at_full_resolution(host, kind):
    """SURVEY.md 8(f) N2, the parity question: the reference triangulates the support points with cv::Subdiv2D
    (ACMMP.cpp:932-954); this library with host/delaunay.cpp.  cv2.Subdiv2D is the same OpenCV implementation, so the
    comparison below pins the stand-in to what the reference would compute, at the size that matters: one support point
    per 5x5 cell of a 3200x2130 view (272 640 points, ~545 000 triangles).  The Delaunay triangulation is unique up to
    co-circular point sets -- common on an integer lattice -- where either diagonal of the quadrilateral is valid; such
    triangles must be the ONLY differences: every triangle one side has and the other lacks is checked (exactly, in
    integers) to have a fourth point ON its circumcircle belonging to the other side's triangle pair."""
    import cv2
    rng = np.random.default_rng(9)
    W, H = 3200, 2130
    cx, cy = np.meshgrid(np.arange(W // 5), np.arange(H // 5), indexing="ij")
    pts = np.stack([5 * cx.ravel() + rng.integers(0, 5, cx.size), 5 * cy.ravel() + rng.integers(0, 5, cx.size)], 1).astype(np.int32)
    if kind == "ragged":
        # real support sets have holes (cells whose best cost is >= 0.1) and a ragged hull: keep 85 % of the cells and cut
        # two corners away
        keep = (rng.random(len(pts)) < 0.85) & (pts[:, 0] + pts[:, 1] > 400) & (pts[:, 0] - pts[:, 1] < 2900)
        pts = np.ascontiguousarray(pts[keep])
    elif kind == "clustered":
        # holes in clusters (texture-less regions) reaching the image border: long thin triangles along the hull, whose
        # existence depends on the virtual enclosing triangle (delaunay.cpp: the one of this image's OpenCV)
        field = cv2.GaussianBlur(rng.random((H // 5, W // 5)).astype(np.float32), (0, 0), 6)
        keep = field[np.minimum(pts[:, 1] // 5, H // 5 - 1), np.minimum(pts[:, 0] // 5, W // 5 - 1)] > np.quantile(field, 0.25)
        pts = np.ascontiguousarray(pts[keep])
    out = np.zeros((2 * len(pts) + 16, 3), np.int32)
    nt = host.acmmp_host_delaunay_rect(pts.ctypes.data_as(C.POINTER(C.c_int32)), len(pts), W, H, out.ctypes.data_as(C.POINTER(C.c_int32)), len(out))
    mine = pts[out[:nt]].astype(np.int64)                                    # [nt, 3, 2]
    sub = cv2.Subdiv2D((0, 0, W, H))
    sub.insert([(float(x), float(y)) for x, y in pts])
    theirs = sub.getTriangleList().astype(np.int64).reshape(-1, 3, 2)
    inside = np.all((theirs[..., 0] >= 0) & (theirs[..., 0] < W) & (theirs[..., 1] >= 0) & (theirs[..., 1] < H), axis=1)
    theirs = theirs[inside]                                                  # main.cpp:140-146 keeps these only

    def keys(t):
        code = t[..., 0] * 4096 + t[..., 1]                                  # one integer per vertex
        code = np.sort(code, axis=1)
        return code[:, 0] * (1 << 50) + code[:, 1] * (1 << 25) + code[:, 2]  # object-free: 3 x 25 bits
    km, kt = keys(mine), keys(theirs)
    common = np.intersect1d(km, kt)
    only_mine = mine[~np.isin(km, common)]
    only_theirs = theirs[~np.isin(kt, common)]
    frac_common = len(common) / max(len(kt), 1)
    # every differing triangle: a fourth scene point lies exactly ON its circumcircle (co-circular tie)
    from scipy.spatial import cKDTree
    tree = cKDTree(pts.astype(np.float64))

    def on_circle_fraction(tris):
        if len(tris) == 0:
            return 1.0
        hit = 0
        for t in tris:
            (ax, ay), (bx, by), (cx_, cy_) = [(int(v[0]), int(v[1])) for v in t]
            centre = t.mean(axis=0)
            near = tree.query_ball_point(centre.astype(np.float64), 16.0)
            found = False
            for k in near:
                dx, dy = int(pts[k][0]), int(pts[k][1])
                if (dx, dy) in ((ax, ay), (bx, by), (cx_, cy_)):
                    continue
                m = [[ax - dx, ay - dy], [bx - dx, by - dy], [cx_ - dx, cy_ - dy]]
                r = [u * u + v * v for u, v in m]
                det = (m[0][0] * (m[1][1] * r[2] - r[1] * m[2][1]) - m[0][1] * (m[1][0] * r[2] - r[1] * m[2][0])
                       + r[0] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]))
                if det == 0:
                    found = True
                    break
            hit += found
        return hit / len(tris)
    sample = rng.permutation(len(only_mine))[:400]
    res = dict(points=len(pts), triangles_mine=int(nt), triangles_subdiv2d_inside=int(len(theirs)), common_fraction=frac_common,
               only_mine=int(len(only_mine)), only_theirs=int(len(only_theirs)),
               only_mine_cocircular=on_circle_fraction(only_mine[sample]),
               only_theirs_cocircular=on_circle_fraction(only_theirs[rng.permutation(len(only_theirs))[:400]]))
    import sys
    sys.path.insert(0, str(ROOT / "tests"))
    import util
    util.dump("delaunay_vs_subdiv2d_c2_" + kind, res)
    # measured: 544 859 triangles on both sides, every one of them in common (the two implementations even break the
    # co-circular ties the same way on this input)
    assert frac_common >= 0.999, res
    assert res["only_theirs_cocircular"] >= 0.98 and res["only_mine_cocircular"] >= 0.98, res


@pytest.mark.parametrize("shape,new", [((205, 410), (160, 320)), ((64, 96), (100, 141)), ((300, 400), (150, 200)), ((33, 57), (33, 29))])
def test_colour_resize_matches_cv_resize_on_8_bit_images(host, shape, new):
    """ResizeLinearBgr = cv::resize INTER_LINEAR on an 8-bit 3-channel image (RescaleImageAndCamera, ACMMP.cpp:236; the
    fusion's colour input): OpenCV's fixed-point scheme.  cv2 wheels may route 8-bit resizes through a SIMD / IPP path
    that rounds differently by one level on a few pixels, hence <= 1 everywhere and equal almost everywhere."""
    import cv2
    rng = np.random.default_rng(4)
    src = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    src[: shape[0] // 2] = np.clip(np.cumsum(rng.integers(-3, 4, (shape[0] // 2, shape[1], 3)), axis=1) + 128, 0, 255).astype(np.uint8)
    nh, nw = new
    dst = np.zeros((nh, nw, 3), np.uint8)
    u8 = C.POINTER(C.c_ubyte)
    assert host.acmmp_host_resize_linear_bgr(src.ctypes.data_as(u8), shape[1], shape[0], dst.ctypes.data_as(u8), nw, nh) == 0
    want = cv2.resize(src, (nw, nh), interpolation=cv2.INTER_LINEAR)
    d = np.abs(dst.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1, d.max()
    assert (d == 0).mean() >= 0.98, (d == 0).mean()


def test_colour_image_loader_reads_ppm_twins_in_opencv_channel_order(host, tmp_path):
    """images/%08d.ppm (P6: R, G, B) -> B, G, R like cv::imread(IMREAD_COLOR); a view with only a .pgm gets (g, g, g)."""
    import cv2
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, (23, 31, 3), dtype=np.uint8)
    (tmp_path / "images").mkdir()
    assert cv2.imwrite(str(tmp_path / "images" / "00000000.ppm"), bgr)                 # cv2 takes B, G, R and writes R, G, B
    grey = rng.integers(0, 256, (23, 31), dtype=np.uint8)
    assert cv2.imwrite(str(tmp_path / "images" / "00000001.pgm"), grey)
    u8 = C.POINTER(C.c_ubyte)
    for view, want in ((0, bgr), (1, np.repeat(grey[..., None], 3, axis=-1))):
        got = np.zeros((23, 31, 3), np.uint8)
        w, h = C.c_int(), C.c_int()
        assert host.acmmp_host_load_colour(str(tmp_path).encode(), view, got.ctypes.data_as(u8), got.size, C.byref(w), C.byref(h)) == 0
        assert (w.value, h.value) == (31, 23) and np.array_equal(got, want)


def test_image_size_comes_from_the_file_header(host, tmp_path):
    """ImageSize (main.cpp:35-71 asks cv::imread for it) reads the PGM header or scans the JPEG for its frame header --
    also behind a large application segment (EXIF with a thumbnail), which lies beyond the first read."""
    import cv2
    rng = np.random.default_rng(7)
    (tmp_path / "images").mkdir()
    cv2.imwrite(str(tmp_path / "images" / "00000000.pgm"), rng.integers(0, 256, (37, 91), dtype=np.uint8))
    cv2.imwrite(str(tmp_path / "images" / "00000001.jpg"), rng.integers(0, 256, (53, 131, 3), dtype=np.uint8))
    ok, enc = cv2.imencode(".jpg", rng.integers(0, 256, (45, 77), dtype=np.uint8))
    raw = enc.tobytes()
    # five APP1 segments of 65 533 payload bytes each between SOI and the rest: the frame header sits past 256 KB
    app1 = b"\xff\xe1" + (65535).to_bytes(2, "big") + bytes(65533)
    (tmp_path / "images" / "00000002.jpg").write_bytes(raw[:2] + app1 * 5 + raw[2:])
    assert cv2.imread(str(tmp_path / "images" / "00000002.jpg"), cv2.IMREAD_GRAYSCALE).shape == (45, 77)
    for view, (w, h) in enumerate([(91, 37), (131, 53), (77, 45)]):
        cw, ch = C.c_int(), C.c_int()
        assert host.acmmp_host_image_size(str(tmp_path).encode(), view, C.byref(cw), C.byref(ch)) == 0
        assert (cw.value, ch.value) == (w, h)
    assert host.acmmp_host_image_size(str(tmp_path).encode(), 9, C.byref(cw), C.byref(ch)) != 0
