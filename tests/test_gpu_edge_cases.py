"""GPU parity at the edges of the input space (through the C ABI, against the compiled reference): image sizes that
are not multiples of any tile, a single source view, the maximum of 32 source views (reference ACMMP.cu:522),
source views of different sizes (each view is scaled to its own size, ACMMP.cpp:606-610), argument errors."""
import ctypes as C

import numpy as np
import pytest

import util
from util import close_frac, dump

pytestmark = pytest.mark.gpu
SEED = 99


def _pair(imgs, cams):
    from acmmp_b200 import Context
    from oracle.ref_driver import RefACMMP
    ctx = Context(0)
    ctx.set_views(imgs, cams)
    ctx.set_seed(SEED)
    return ctx, RefACMMP(imgs, cams, seed=SEED)


def _stage_agreement(ctx, ref, gt):
    """Whole stage on both sides.  The reference races and carries an uninitialised variable, so whole
    maps agree statistically; what must hold at any size: the same quality against ground truth, the same share of
    NaN costs (pixels whose 15 view draws all fail divide 0 by 0 in the reference too, ACMMP.cu:1243), no inf."""
    ctx.run_patch_match()
    pa, ca = ctx.get_result()
    ref.run_patch_match()
    pb, cb = ref.get_result()
    rel = np.abs(pa[..., 3] - pb[..., 3]) / np.maximum(np.abs(pb[..., 3]), 1e-9)
    return dict(depth_within_1pct=float((rel <= 0.01).mean()),
                mine_vs_gt=float((np.abs(pa[..., 3] - gt) / gt <= 0.01).mean()),
                ref_vs_gt=float((np.abs(pb[..., 3] - gt) / gt <= 0.01).mean()),
                nan_mine=float(np.isnan(ca).mean()), nan_ref=float(np.isnan(cb).mean()),
                no_inf=bool(not np.isinf(pa).any() and not np.isinf(ca).any()))


def _one_pass(ctx, ref, H, W):
    """Deterministic comparison: one black pass from the reference's own post-init state (test_gpu_parity._pass_compare)."""
    from test_gpu_parity import _pass_compare
    ref.launch_init()
    r, _, _ = _pass_compare(ctx, ref, 0, 0, H, W, border=min(4, H // 8))
    return max(r["mode0"]["plane_match"], r["mode1"]["plane_match"]), r["mode1"]["untouched_ok"], max(r["mode0"]["rng_match"], r["mode1"]["rng_match"])


def _quality_ok(res, slack=0.06, nan_slack=0.03):
    """(whole-stage figures are statistical, see _stage_agreement; on a 40 x 24 image 1 % is ten pixels)"""
    return res["no_inf"] and abs(res["mine_vs_gt"] - res["ref_vs_gt"]) <= slack and abs(res["nan_mine"] - res["nan_ref"]) <= nan_slack


@pytest.mark.parametrize("size", [(333, 251), (97, 61), (40, 24)])
def test_ragged_image_sizes(size):
    """Sizes that no tile divides (8x16 pass tiles, 16x8 init tiles, odd widths for the checkerboard)."""
    from acmmp_b200 import synth
    w, h = size
    scene = synth.make_pinhole_scene(n_views=4, width=w, height=h, focal=0.8 * w, seed=11)
    imgs, cams, ids = scene.problem(0)
    ctx, ref = _pair(imgs, cams)
    planes = util.random_planes(scene, 0, seed=3, perturb=0.05)
    a, b = ctx.probe_ncc(planes, 1), ref.probe_ncc(planes, 1)
    res = dict(ncc_1e4=close_frac(a, b, atol=1e-4, rtol=1e-4), ncc_1e3=close_frac(a, b, atol=1e-3, rtol=1e-3))
    # identical initial hypotheses (same XORWOW stream per pixel) everywhere, borders included
    ctx.random_init(); ctx.synchronize()
    ref.launch_init()
    sa, sb = ctx.download_state(), ref.download_state()
    res["init_planes"] = close_frac(sa["planes"], sb["planes"], 1e-5, 1e-5)
    res["init_views"] = float((sa["views"] == sb["views"]).mean())
    res["pass_plane_match"], res["pass_untouched"], res["pass_rng"] = _one_pass(ctx, ref, h, w)
    res.update(_stage_agreement(ctx, ref, scene.depths_gt[0]))
    dump(f"edge_size_{w}x{h}", res)
    assert res["ncc_1e4"] >= 0.97 and res["ncc_1e3"] >= 0.999, res
    assert res["init_planes"] >= 0.999 and res["init_views"] >= 0.99, res
    assert res["pass_plane_match"] >= 0.90 and res["pass_untouched"] == 1.0 and res["pass_rng"] >= 0.97, res
    assert _quality_ok(res), res


def test_single_source_view():
    from acmmp_b200 import synth
    scene = synth.make_pinhole_scene(n_views=2, width=160, height=120, focal=130.0, seed=5)
    imgs, cams, ids = scene.problem(0)
    assert len(imgs) == 2
    ctx, ref = _pair(imgs, cams)
    res = {}
    res["pass_plane_match"], res["pass_untouched"], res["pass_rng"] = _one_pass(ctx, ref, 120, 160)
    res.update(_stage_agreement(ctx, ref, scene.depths_gt[0]))
    dump("edge_single_source", res)
    assert res["pass_plane_match"] >= 0.90 and res["pass_untouched"] == 1.0, res
    assert _quality_ok(res), res


def test_thirty_two_source_views():
    """The reference's arrays and view bitmask hold at most 32 source views (ACMMP.cu:522, :88-96)."""
    from acmmp_b200 import synth, AcmmpError, Context
    scene = synth.make_pinhole_scene(n_views=34, width=128, height=96, focal=110.0, seed=6, baseline_ratio=0.012, n_src=33)
    imgs, cams, ids = scene.problem(16)
    imgs, cams = imgs[:33], cams[:33]            # reference view + 32 sources
    ctx, ref = _pair(imgs, cams)
    planes = util.random_planes(scene, 16, seed=3, perturb=0.02)
    a, va = ctx.probe_initcost(planes)
    b, vb = ref.probe_initcost(planes)
    res = dict(initcost=close_frac(a, b, atol=1e-3, rtol=1e-3), views=float((va == vb).mean()), bit31_used=bool((va >> 31).any()))
    res["pass_plane_match"], res["pass_untouched"], res["pass_rng"] = _one_pass(ctx, ref, 96, 128)
    res.update(_stage_agreement(ctx, ref, scene.depths_gt[16]))
    dump("edge_32_sources", res)
    assert res["initcost"] >= 0.995 and res["views"] >= 0.98, res
    assert res["pass_plane_match"] >= 0.90 and res["pass_untouched"] == 1.0, res
    assert _quality_ok(res), res
    # 33 source views are refused, not truncated
    imgs34, cams34, _ = scene.problem(16)
    with pytest.raises(AcmmpError):
        Context(0).set_views(imgs34[:34], cams34[:34])


def test_source_views_of_different_sizes():
    """Every view is scaled to its own size upstream; sources smaller than the layered texture are edge-replicated."""
    import cv2
    from acmmp_b200 import synth, make_camera, MODEL_PINHOLE
    scene = synth.make_pinhole_scene(n_views=4, width=200, height=150, focal=170.0, seed=8)
    imgs, cams, ids = scene.problem(0)
    imgs, cams = list(imgs), list(cams)
    # halve source view 2 (image + intrinsics), as RescaleImageAndCamera would (ACMMP.cpp:225-246)
    small = cv2.resize(imgs[2], (100, 75), interpolation=cv2.INTER_LINEAR)
    K = np.array(list(cams[2].K), np.float32).reshape(3, 3).copy()
    K[0, 0] *= 0.5; K[0, 2] *= 0.5; K[1, 1] *= 0.5; K[1, 2] *= 0.5
    cams[2] = make_camera(MODEL_PINHOLE, list(cams[2].R), list(cams[2].t), K=K, width=100, height=75,
                          depth_min=cams[2].depth_min, depth_max=cams[2].depth_max)
    imgs[2] = small
    ctx, ref = _pair(imgs, cams)
    planes = util.random_planes(scene, 0, seed=3, perturb=0.03)
    res = {}
    for view in (1, 2, 3):
        a, b = ctx.probe_ncc(planes, view), ref.probe_ncc(planes, view)
        res[f"ncc_v{view}_1e4"] = close_frac(a, b, atol=1e-4, rtol=1e-4)
        res[f"ncc_v{view}_1e3"] = close_frac(a, b, atol=1e-3, rtol=1e-3)
    res["pass_plane_match"], res["pass_untouched"], res["pass_rng"] = _one_pass(ctx, ref, 150, 200)
    res.update(_stage_agreement(ctx, ref, scene.depths_gt[0]))
    dump("edge_mixed_sizes", res)
    for view in (1, 2, 3):
        assert res[f"ncc_v{view}_1e4"] >= 0.97 and res[f"ncc_v{view}_1e3"] >= 0.999, res
    assert res["pass_plane_match"] >= 0.90 and res["pass_untouched"] == 1.0, res
    assert _quality_ok(res), res


def test_argument_errors_are_reported_not_fatal():
    """The reference exit()s on failure (ACMMP.cpp:64-97); the library returns codes."""
    from acmmp_b200 import Context, AcmmpError, lib, synth
    ctx = Context(0)
    with pytest.raises(AcmmpError):
        ctx.run_patch_match()                      # no views yet
    scene = synth.make_pinhole_scene(n_views=3, width=64, height=48, focal=60.0, seed=1)
    imgs, cams, ids = scene.problem(0)
    with pytest.raises(AcmmpError):
        ctx.set_views(imgs[:1], cams[:1])          # a reference view without sources
    ctx.set_views(imgs, cams)
    ctx.set_geom_consistency(False)
    with pytest.raises(AcmmpError):
        ctx.run_patch_match()                      # geometric mode without depth maps
    ctx.reset_modes()
    ctx.set_planar_prior()
    with pytest.raises(AcmmpError):
        ctx.run_patch_match()                      # prior mode without prior inputs
    ctx.reset_modes()
    ctx.run_patch_match()                          # and the context is still usable
    assert lib().acmmp_create(C.byref(C.c_void_p()), C.c_int(4096)) != 0      # no such device


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_device_planar_prior_against_the_numpy_twin(model):
    """acmmp_support_points + acmmp_planar_prior_from_triangles (k_support_cells, k_tri_planes,
    k_tri_raster[_long], k_prior_finish) against acmmp_b200/prior.py, which follows the reference's CPU stage
    (ACMMP.cpp:904-1011, main.cpp:113-185) with cv2.Subdiv2D -- the reference's own triangulator -- on the SAME photometric
    result: same support points, same triangles (tests/test_cpu_cpp_host.py pins the C++ triangulator to Subdiv2D), same
    plane per triangle; per pixel the same prior plane except on triangle edges, where the twin's cv2.fillConvexPoly and the
    reference's stepping rasteriser (reproduced on the device) assign the pixel to either neighbour."""
    import cv2
    from acmmp_b200 import Context, synth
    from acmmp_b200 import prior as twin
    from acmmp_b200.scene import delaunay_triangles_inside
    scene = (synth.make_pinhole_scene(n_views=4, width=640, height=480, focal=500.0, seed=1) if model == "pinhole"
             else synth.make_sphere_scene(n_views=4, width=1024, height=512, seed=4))
    imgs, cams, _ = scene.problem(0)
    ctx = Context(0)
    ctx.set_views(imgs, cams)
    ctx.set_seed(5)
    ctx.run_patch_match()
    planes, costs = (np.array(a) for a in ctx.get_result())
    H, W = costs.shape
    dmin = float(np.float32(cams[0].depth_min) * np.float32(0.6))
    dmax = float(np.float32(cams[0].depth_max) * np.float32(1.2))
    ctx.set_planar_prior()
    pts = ctx.support_points()
    pts_twin = twin.support_points(costs)
    assert np.array_equal(pts, pts_twin)
    ctx.planar_prior_from_triangles(delaunay_triangles_inside(pts, W, H))
    pp, mk = ctx.download_prior()
    p_ref, m_ref = twin.planar_prior(cams[0], np.ascontiguousarray(planes[..., 3]), costs, dmin, dmax)
    both = (mk > 0) & (m_ref > 0)
    mine = pp[both]
    theirs = p_ref[m_ref[both].astype(np.int64) - 1]
    same = np.all(np.abs(mine - theirs) <= 1e-4 + 1e-4 * np.abs(theirs), axis=1)
    # the planes of the pixels that disagree belong to a NEIGHBOURING triangle of the twin: check against the twin's plane set
    from scipy.spatial import cKDTree
    scale = np.abs(p_ref).max(axis=0)
    d, _ = cKDTree(p_ref / scale).query(mine[~same] / scale) if (~same).any() else (np.zeros(0), None)
    res = dict(support_points=int(len(pts)), triangles=int(mk.max()), triangles_twin=int(len(p_ref)),
               masked_mine=float((mk > 0).mean()), masked_twin=float((m_ref > 0).mean()), masked_both=float(both.mean()),
               same_plane_where_both=float(same.mean()), disagreeing_planes_found_in_twin_set=float((d < 1e-4).mean()) if len(d) else 1.0)
    util.dump(f"device_prior_vs_twin_{model}", res)
    ctx.close()
    assert res["triangles"] == res["triangles_twin"], res
    assert abs(res["masked_mine"] - res["masked_twin"]) < 0.02 and res["masked_both"] > 0.9 * res["masked_twin"], res
    assert res["same_plane_where_both"] > 0.6, res
    assert res["disagreeing_planes_found_in_twin_set"] > 0.99, res


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("stage", ["geom", "prior", "hierarchy_prior"])
def test_parked_context_resumes_bit_identically(stage):
    """acmmp_park between two stages of a resident view (the scratch goes to the device's pool, another context with the
    same seed and shapes runs in it, the view is re-activated with acmmp_set_views of the same shapes): the second stage
    gives the bits it gives on a context that never let go of anything.  hierarchy_prior: the prior stage of the SECOND
    pyramid level, whose passes read what the photometric stage's initialisation left (pre_costs, ACMMP.cu:1318-1322)."""
    from acmmp_b200 import Context, synth
    from acmmp_b200.scene import delaunay_triangles_inside
    scene = synth.make_pinhole_scene(n_views=4, width=640, height=480, focal=520.0, seed=5)
    full, full_cams, ids = scene.problem(0)
    other_full, other_full_cams, _ = scene.problem(1)
    imgs, cams = synth.scale_problem(full, full_cams, 320)
    other_imgs, other_cams = synth.scale_problem(other_full, other_full_cams, 320)
    two_levels = stage == "hierarchy_prior"

    def first_stage(ctx, im0, cm0, im1, cm1):
        ctx.set_seed(SEED)
        ctx.set_views(im0, cm0)
        ctx.run_patch_match(download=False)
        if two_levels:
            ctx.next_level(im1, cm1)
            ctx.run_patch_match(download=False)

    def second_stage(ctx, pts=None):
        H, W = ctx.H, ctx.W
        if stage == "geom":
            import cv2
            ctx.reset_modes()
            ctx.set_geom_consistency(False)
            ctx.set_depth_maps([None] + [cv2.resize(scene.depths_gt[i], (W, H), interpolation=cv2.INTER_NEAREST) for i in ids[1:]])
        else:
            ctx.set_planar_prior()
            pts = ctx.support_points() if pts is None else pts
            assert len(pts) > 50
            ctx.planar_prior_from_triangles(delaunay_triangles_inside(pts, W, H))
        ctx.run_patch_match()
        return ctx.get_result()

    plain = Context(0)
    first_stage(plain, imgs, cams, full, full_cams)
    p0, c0 = second_stage(plain)
    plain.close()

    a = Context(0)
    first_stage(a, imgs, cams, full, full_cams)
    a.park()
    with pytest.raises(RuntimeError):
        a.run_patch_match(download=False)                       # parked: no scratch to run in
    pts = a.support_points() if stage != "geom" else None       # reads the state only: works on a parked context
    b = Context(0)                                              # another view takes the blocks a let go of
    first_stage(b, other_imgs, other_cams, other_full, other_full_cams)
    b.park()
    a.set_views(full if two_levels else imgs, full_cams if two_levels else cams)      # same shapes: the state is kept
    p1, c1 = second_stage(a, pts)
    a.close()
    b.close()
    assert np.array_equal(_bits(p0), _bits(p1)) and np.array_equal(_bits(c0), _bits(c1))


def test_reserved_device_memory_serves_contexts_and_host_buffers():
    """acmmp_reserve_device_memory: the device's pool carves blocks out of one reservation; results do not depend on where a
    buffer lives; acmmp_pool_alloc hands the host buffers out of the same pool; a second reservation is refused while the
    first one is alive; everything is returned when the last context / host block is gone."""
    import gc
    import acmmp_b200 as ab
    from acmmp_b200 import Context, synth
    gc.collect()                                                   # contexts other tests dropped without close()
    scene = synth.make_pinhole_scene(n_views=3, width=200, height=150, focal=160.0, seed=6)
    imgs, cams, _ = scene.problem(0)

    def stage():
        ctx = Context(0)
        ctx.set_seed(SEED)
        ctx.set_views(imgs, cams)
        ctx.run_patch_match()
        out = ctx.get_result()
        return ctx, out

    ctx, (p0, c0) = stage()
    ctx.close()                                                    # last context of the device: the pool is empty again
    assert ab.reserve_device_memory(0, 256 << 20) == 0
    assert ab.reserve_device_memory(0, 1 << 20) == -1              # ACMMP_E_ARG: one reservation per device
    ctx, (p1, c1) = stage()
    planes_dev, _ = ctx.device_buffers()
    host_block = ab.pool_alloc(0, 4 * 200 * 150)
    ctx.export_depth_device(host_block)
    ctx.synchronize()
    ctx.close()                                                    # the host block keeps the pool (and the reservation) alive
    assert ab.reserve_device_memory(0, 1 << 20) == -1
    again = ab.pool_alloc(0, 4 * 200 * 150)
    assert again != host_block
    # both blocks lie inside one 256 MB range: carved out of the reservation, not separate allocations
    assert abs(again - host_block) < (256 << 20) and abs(int(planes_dev) - host_block) < (256 << 20)
    ab.pool_free(0, again)
    ab.pool_free(0, host_block)                                    # last reference: reservation released
    assert ab.reserve_device_memory(0, 1 << 20) == 0
    ctx, _ = stage()
    ctx.close()
    assert np.array_equal(_bits(p0), _bits(p1)) and np.array_equal(_bits(c0), _bits(c1))
