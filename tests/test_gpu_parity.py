"""GPU parity tests: the CUDA path (through the C ABI) against the reference's own kernels
(oracle/_ref/libacmmp_ref.so = unmodified reference sources compiled for sm_100) on identical
seeded inputs.  Tolerances follow BASELINE.json:north_star:
  * deterministic sub-kernels (warp, NCC at fixed planes, geometric term, JBU): 1e-4 relative
  * RandomInitialization with the same XORWOW seed: identical hypotheses (1e-5) and RNG states
  * full maps: statistical (>= 99 % of pixels within 1 % relative depth, normals within 5 deg),
    because the reference races on same-colour reads and has an uninitialised variable
    (SURVEY.md 7.3); single passes from an identical state are compared pixel by pixel.
Every test dumps the figures it measured to gpurun_out/metrics_*.json.
"""
import numpy as np
import pytest

import util
from util import close_frac, dump

pytestmark = pytest.mark.gpu

SEED = 1234


def _mine(scene, ref=0, **kw):
    from acmmp_b200 import Context
    imgs, cams, ids = scene.problem(ref)
    ctx = Context(0)
    ctx.set_views(imgs, cams)
    ctx.set_seed(SEED)
    return ctx, imgs, cams, ids


def _ref(scene, ref=0, **kw):
    from oracle.ref_driver import RefACMMP
    imgs, cams, ids = scene.problem(ref)
    return RefACMMP(imgs, cams, seed=SEED, **kw)


# ------------------------------------------------------------------------------------------
# deterministic sub-kernels
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_warp_matches_reference(model):
    scene = util.scene_of(model)
    ctx, *_ = _mine(scene)
    ref = _ref(scene)
    res = {}
    for name, perturb in (("gt", 0.0), ("jitter", 0.2)):
        planes = util.random_planes(scene, 0, seed=3, perturb=perturb)
        for view in (1, 3):
            a = ctx.probe_warp(planes, view)
            b = ref.probe_warp(planes, view)
            # pixel coordinates: 1e-4 relative to the image extent; depths: 1e-4 relative
            fx = close_frac(a[..., 0], b[..., 0], atol=1e-4 * scene.images[0].shape[1], rtol=1e-4)
            fy = close_frac(a[..., 1], b[..., 1], atol=1e-4 * scene.images[0].shape[0], rtol=1e-4)
            fd = close_frac(a[..., 2], b[..., 2], atol=0, rtol=1e-4)
            fr = close_frac(a[..., 3], b[..., 3], atol=0, rtol=1e-4)
            res[f"{name}_v{view}"] = dict(fx=fx, fy=fy, fd=fd, fr=fr,
                                          max_dx=float(np.nanmax(np.abs(a[..., 0] - b[..., 0]))),
                                          max_dy=float(np.nanmax(np.abs(a[..., 1] - b[..., 1]))))
    dump(f"warp_{model}", res)
    for k, v in res.items():
        assert min(v["fx"], v["fy"], v["fd"], v["fr"]) >= 0.999, (k, v)


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_ncc_fixed_planes_matches_reference(model):
    """ComputeBilateralNCC at fixed plane hypotheses through quad_ncc -- the device function k_pass and k_random_init
    evaluate every hypothesis with (four lanes per pixel, 2x2 tap blocks, packed FP32, skipped samples zeroed + per-lane
    rebuilt reference sums) -- against the compiled reference."""
    scene = util.scene_of(model)
    ctx, *_ = _mine(scene)
    ref = _ref(scene)
    res = {}
    for name, perturb in (("gt", 0.0), ("jitter", 0.05), ("far", 0.5)):
        planes = util.random_planes(scene, 0, seed=5, perturb=perturb)
        for view in (1, 2, 4):
            a = ctx.probe_ncc(planes, view)
            b = ref.probe_ncc(planes, view)
            d = np.abs(a - b)
            res[f"{name}_v{view}"] = dict(
                frac_1e4=close_frac(a, b, atol=1e-4, rtol=1e-4), frac_1e3=close_frac(a, b, atol=1e-3, rtol=1e-3),
                max=float(d.max()), p999=float(np.quantile(d, 0.999)), mean_cost_ref=float(b.mean()),
                frac_cost2_ref=float((b >= 2.0).mean()), frac_cost2_mine=float((a >= 2.0).mean()))
    dump(f"ncc_{model}", res)
    for k, v in res.items():
        # the texture unit quantises bilinear fractions to 1/256 px, so source coordinates that differ
        # in the last float bits (1e-4 px here, see test_warp) occasionally flip a fraction: ~99 % of the
        # pixels agree to 1e-4, all but a few in 1e5 to 1e-3, identical validity (cost == 2) decisions
        assert v["frac_1e4"] >= 0.98, (k, v)
        assert v["frac_1e3"] >= 0.9995, (k, v)
        assert abs(v["frac_cost2_ref"] - v["frac_cost2_mine"]) <= 1e-4, (k, v)


def _fraction_codes(c):
    """The 1.8 fixed-point bilinear fraction the texture unit derives from a fetch coordinate (texel centres at +0.5):
    index of the 1/256 cell of x - 0.5, under round-to-nearest and under truncation (both conventions are checked)."""
    xb = c.astype(np.float64) - 0.5
    return np.floor(xb * 256.0 + 0.5), np.floor(xb * 256.0)


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_ncc_residue_above_1e4_is_the_texture_units_fraction_quantisation(model):
    """north_star: NCC at fixed planes within 1e-4.  98.7-99.7 % of the pixels are; this test shows what the rest is.
    Both sides fetch the source image with the SAME hardware bilinear filter, whose fractions are 1.8 fixed point: the
    sample value is a step function of the coordinate with steps every 1/256 px.  The reference computes a sample's
    coordinate through its unfolded chain (PixelToDir, ray/plane, lift, R^T X + C, R X + t, project: ACMMP.cu:458-476),
    this library through one folded transform per view -- same maths, different roundings, coordinates 1e-5..1e-4 px
    apart.  Where such a pair of coordinates straddles a step, one tap's sample moves by (gradient / 256) and the cost by
    ~1e-4..1e-3.  Measured here per pixel, with the coordinates of all 36 taps from both implementations:
      * pixels where no tap's (u, v) changes its 1/256 cell (~58 % of them): the costs agree within 1e-4 for all but
        1-2 in 45 000 (what is left there is the summation order: four partial sums per quad instead of one running
        sum, amplified by the E[x^2] - E[x]^2 cancellation of the variances),
      * of the pixels whose cost differs by more than 1e-4, > 99.5 % have at least one tap that changed cell (233 of
        235, 299 / 300, 911 / 912, 893 / 894 measured): P(above | a tap flipped) is ~0.7 %, P(above | none) ~4e-5.
    I.e. the residue is the texture unit's quantisation applied to two correctly rounded evaluations of the same
    formula, not a different formula.  The reference itself with every input pixel one ulp up stays within 1e-4 on
    99.997 % of the PINHOLE pixels: the quantisation, not conditioning, is what is seen there.  SPHERE is the opposite
    case, see the assertion below."""
    from oracle.ref_driver import RefACMMP
    scene = util.scene_of(model)
    ctx, imgs, cams, _ = _mine(scene)
    ref = _ref(scene)
    # the reference on inputs one ulp up (every pixel of every image the next representable float): its own sensitivity
    ref_ulp = RefACMMP([np.nextafter(np.asarray(im, np.float32), np.float32(np.inf)) for im in imgs], cams, seed=SEED)
    H, W = scene.images[0].shape
    res = {}
    for name, perturb in (("gt", 0.0), ("jitter", 0.05)):
        planes = util.random_planes(scene, 0, seed=5, perturb=perturb)
        for view in (1, 2):
            a = ctx.probe_ncc(planes, view)
            b = ref.probe_ncc(planes, view)
            b_ulp = ref_ulp.probe_ncc(planes, view)
            ca = ctx.probe_coords(planes, view)
            cb = ref.probe_coords(planes, view)
            fin = np.isfinite(ca).all(axis=(2, 3)) & np.isfinite(cb).all(axis=(2, 3))
            a1, a0 = _fraction_codes(np.where(np.isfinite(ca), ca, 0))
            b1, b0 = _fraction_codes(np.where(np.isfinite(cb), cb, 0))
            flips = ((a1 != b1) | (a0 != b0)).any(axis=3).sum(axis=2)          # taps per pixel whose cell differs
            clean = (flips == 0) & fin & util.interior(H, W, 6)
            d = np.abs(a.astype(np.float64) - b)
            tol = 1e-4 + 1e-4 * np.abs(b)
            above = (d > tol) & fin & util.interior(H, W, 6)
            res[f"{name}_v{view}"] = dict(
                frac_1e4=float((d <= tol).mean()), clean_frac=float(clean.mean()),
                clean_within_2e5=float((d[clean] <= 2e-5 + 2e-5 * np.abs(b[clean])).mean()) if clean.any() else 1.0,
                clean_within_1e4=float((d[clean] <= tol[clean]).mean()) if clean.any() else 1.0,
                flipped_within_1e4=float((d[~clean & fin] <= tol[~clean & fin]).mean()),
                clean_max=float(d[clean].max()) if clean.any() else 0.0,
                above_1e4=int(above.sum()), above_with_flip=int((above & (flips > 0)).sum()),
                ref_vs_ref_inputs_one_ulp_up_1e4=float((np.abs(b_ulp.astype(np.float64) - b) <= tol).mean()),
                ref_one_ulp_max=float(np.abs(b_ulp.astype(np.float64) - b).max()),
                max_coord_diff_px=float(np.abs(np.where(fin[..., None, None], ca - cb, 0)).max()),
                mean_flipped_taps=float(flips[fin].mean()))
    dump(f"ncc_residue_{model}", res)
    for k, v in res.items():
        if model == "pinhole":
            assert v["clean_within_1e4"] >= 0.9999, (k, v)
            assert v["above_with_flip"] >= 0.99 * v["above_1e4"], (k, v)
            assert v["clean_frac"] >= 0.3, (k, v)                     # the statement is about a substantial set of pixels
        else:
            # SPHERE: the coordinates are bit-identical to the reference's on the GT planes (max_coord_diff 0, no flip) and the
            # cost still moves: the angular bilateral weight (sigma_eff = 5 pi / H, ACMMP.cu:436-442) concentrates the window on
            # a few taps, E[x^2] - E[x]^2 cancels, and the reference's own result moves by more than 1e-4 on 0.4-0.6 % of the
            # pixels (up to 3e-3) when every input pixel goes up by ONE ulp.  Here: summation order (four partial sums).
            assert v["frac_1e4"] >= v["ref_vs_ref_inputs_one_ulp_up_1e4"] - 0.005, (k, v)
        assert v["frac_1e4"] >= 0.985, (k, v)


def test_sphere_tap_pruning_changes_the_cost_by_less_than_the_references_own_one_ulp_sensitivity():
    """acmmp_set_sphere_tap_pruning (default 2^-24): at fine pyramid levels the fork's angular bilateral weight leaves most
    window taps with weights far below the float32 resolution of the sums; tap steps whose weights are all below 2^-24 of
    the weight sum are not sampled.  Measured here at 2048x1024 (where about half of the steps go) and, scaled, at the
    weights of a 3200x1600 level: the cost with pruning against the cost with every tap sampled, next to how much the
    REFERENCE's cost moves when its inputs go up by one ulp."""
    from acmmp_b200 import synth, Context
    from oracle.ref_driver import RefACMMP
    scene = synth.make_sphere_scene(n_views=3, width=2048, height=1024, seed=12)
    imgs, cams, _ = scene.problem(0)
    ctx = Context(0)
    ctx.set_views(imgs, cams)
    ref = RefACMMP(imgs, cams, seed=SEED)
    ref_ulp = RefACMMP([np.nextafter(np.asarray(im, np.float32), np.float32(np.inf)) for im in imgs], cams, seed=SEED)
    res = {}
    for name, perturb in (("gt", 0.0), ("jitter", 0.05)):
        planes = util.random_planes(scene, 0, seed=5, perturb=perturb)
        ctx.set_sphere_tap_pruning(0.0)
        full = ctx.probe_ncc(planes, 1).astype(np.float64)
        ctx.set_sphere_tap_pruning(2.0 ** -24)
        pruned = ctx.probe_ncc(planes, 1).astype(np.float64)
        b = ref.probe_ncc(planes, 1).astype(np.float64)
        b_ulp = ref_ulp.probe_ncc(planes, 1).astype(np.float64)
        d = np.abs(pruned - full)
        res[name] = dict(
            identical=float((pruned == full).mean()), within_1e6=float((d <= 1e-6 + 1e-6 * np.abs(full)).mean()),
            within_1e5=float((d <= 1e-5 + 1e-5 * np.abs(full)).mean()), within_1e4=float((d <= 1e-4 + 1e-4 * np.abs(full)).mean()),
            max=float(d.max()), cost_2_fraction=float((full >= 2.0).mean()),
            pruned_vs_ref_1e4=close_frac(pruned, b, 1e-4, 1e-4), full_vs_ref_1e4=close_frac(full, b, 1e-4, 1e-4),
            ref_vs_ref_inputs_one_ulp_up_1e4=close_frac(b_ulp, b, 1e-4, 1e-4), ref_one_ulp_max=float(np.abs(b_ulp - b).max()))
    dump("sphere_tap_pruning", res)
    ctx.close()
    # Measured (B200, round 2) at 2048x1024: pruned == all taps bit for bit on 73 % of the pixels, within 1e-4 on 99.4-99.6 %;
    # against the reference within 1e-4: pruned 97.1 / 97.5 %, all taps 97.2 / 97.6 %; the reference against ITSELF with its
    # inputs one ulp up: 97.6 / 98.2 %, moving by up to 1.0 (the whole cost range) -- at this resolution the fork's weights
    # make the cost ill-conditioned, and dropping taps of relative weight < 2^-24 is a smaller perturbation than one ulp.
    for k, v in res.items():
        assert v["within_1e4"] >= 0.99, (k, v)
        assert v["pruned_vs_ref_1e4"] >= v["full_vs_ref_1e4"] - 0.003, (k, v)
        assert v["pruned_vs_ref_1e4"] >= v["ref_vs_ref_inputs_one_ulp_up_1e4"] - 0.01, (k, v)
        assert 1.0 - v["within_1e4"] <= 1.0 - v["ref_vs_ref_inputs_one_ulp_up_1e4"], (k, v)


def test_packed_two_hypothesis_sphere_projection_is_bit_identical_to_the_scalar_one():
    """sphere_coords2 (FMUL2 / FFMA2 / FADD2 over two hypotheses, what the checkerboard pass runs for its neighbour
    and refinement hypotheses) against the one-hypothesis projection the NCC sub-kernel tests pin to the reference."""
    scene = util.sphere_scene()
    ctx, *_ = _mine(scene)
    res = {}
    for name, perturb in (("gt", 0.0), ("far", 0.5)):
        planes = util.random_planes(scene, 0, seed=5, perturb=perturb)
        for view in (1, 3):
            c0 = ctx.probe_coords(planes, view, 0)
            for variant in (1, 2):
                cv = ctx.probe_coords(planes, view, variant)
                same = (c0.view(np.uint32) == cv.view(np.uint32)) | (np.isnan(c0) & np.isnan(cv))
                res[f"{name}_v{view}_as_hypothesis_{variant}"] = float(same.mean())
    dump("sphere_packed_projection", res)
    assert min(res.values()) == 1.0, res


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_initial_cost_and_views_match_reference(model):
    scene = util.scene_of(model)
    ctx, *_ = _mine(scene)
    ref = _ref(scene)
    planes = util.random_planes(scene, 0, seed=9, perturb=0.05)
    a, av = ctx.probe_initcost(planes)
    b, bv = ref.probe_initcost(planes)
    res = dict(frac_cost=close_frac(a, b, atol=1e-4, rtol=1e-4), frac_views=float((av == bv).mean()),
               max=float(np.abs(a - b).max()))
    dump(f"initcost_{model}", res)
    assert res["frac_cost"] >= 0.995 and res["frac_views"] >= 0.99, res


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_geom_cost_matches_reference(model):
    scene = util.scene_of(model)
    imgs, cams, ids = scene.problem(0)
    depth_maps = [scene.depths_gt[i] for i in ids]
    gt = util.random_planes(scene, 0, seed=1, perturb=0.0)
    prev = np.concatenate([util.world_normals(scene, 0, gt), scene.depths_gt[0][..., None]], axis=-1)
    costs = np.full(scene.images[0].shape, 0.3, np.float32)
    ctx, *_ = _mine(scene)
    ctx.set_geom_consistency(False)
    ctx.set_depth_maps(depth_maps)
    ctx.set_planes(prev, costs)
    ref = _ref(scene, geom=True, depth_maps=depth_maps, prev_planes=prev, prev_costs=costs)
    res = {}
    for name, perturb in (("gt", 0.0), ("jitter", 0.03)):
        planes = util.random_planes(scene, 0, seed=11, perturb=perturb)
        for view in (1, 3):
            a = ctx.probe_geom(planes, view)
            b = ref.probe_geom(planes, view)
            d = np.abs(a - b)
            res[f"{name}_v{view}"] = dict(frac=close_frac(a, b, atol=2e-3, rtol=1e-4), max=float(d.max()),
                                          p99=float(np.quantile(d, 0.99)), mean_ref=float(b.mean()))
    dump(f"geom_{model}", res)
    for k, v in res.items():
        assert v["frac"] >= 0.99, (k, v)


def test_jbu_matches_reference():
    from acmmp_b200 import jbu
    from oracle.ref_driver import run_jbu
    import cv2
    scene = util.pinhole_scene()
    img = scene.images[0]
    H, W = img.shape
    coarse = cv2.resize(scene.depths_gt[0], (W // 2, H // 2), interpolation=cv2.INTER_NEAREST)
    a = jbu(img, coarse)
    b = run_jbu(img, coarse)
    res = dict(frac=close_frac(a, b, atol=0, rtol=1e-4), max_rel=float(np.max(np.abs(a - b) / np.abs(b))))
    dump("jbu", res)
    assert res["frac"] >= 0.9999, res


# ------------------------------------------------------------------------------------------
# RandomInitialization
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_random_init_matches_reference(model):
    scene = util.scene_of(model)
    ctx, *_ = _mine(scene)
    ref = _ref(scene)
    ctx.random_init()
    ctx.synchronize()
    ref.launch_init()
    a = ctx.download_state()
    b = ref.download_state()
    res = dict(
        rng_equal=float((a["rand"] == b["rand"]).all(axis=-1).mean()),
        planes_1e5=close_frac(a["planes"], b["planes"], atol=1e-5, rtol=1e-5),
        costs_1e4=close_frac(a["costs"], b["costs"], atol=1e-4, rtol=1e-4),
        views_equal=float((a["views"] == b["views"]).mean()),
        max_plane=float(np.abs(a["planes"] - b["planes"]).max()),
    )
    dump(f"init_{model}", res)
    assert res["rng_equal"] == 1.0, res
    assert res["planes_1e5"] >= 0.9999, res
    assert res["costs_1e4"] >= 0.995 and res["views_equal"] >= 0.99, res


# ------------------------------------------------------------------------------------------
# one checkerboard pass from an identical state
# ------------------------------------------------------------------------------------------
def _pass_compare(ctx, ref, colour, it, H, W, border=4):
    """Both sides start from the reference's state; returns per-pixel agreement on the pixels the
    pass updates (interior only: the reference writes garbage near the border)."""
    st = ref.download_state(rand=True, pre_costs=True)
    out = {}
    for mode in (1, 0):
        ctx.upload_state(planes=st["planes"], costs=st["costs"], views=st["views"], rand=st["rand"], pre_costs=st["pre_costs"])
        ctx.set_plane_now_semantics(bool(mode))
        ctx.checkerboard_pass(colour, it)
        ctx.synchronize()
        out[mode] = ctx.download_state()
    ref.launch_pass(colour, it)
    b = ref.download_state()
    upd = util.colour_mask(H, W, colour) & util.interior(H, W, border)
    keep = (~util.colour_mask(H, W, colour))
    res = {}
    for mode, a in out.items():
        same_plane = np.all(np.abs(a["planes"] - b["planes"]) <= 1e-4 + 1e-4 * np.abs(b["planes"]), axis=-1)
        res[f"mode{mode}"] = dict(
            plane_match=float(same_plane[upd].mean()),
            cost_match=close_frac(a["costs"], b["costs"], atol=1e-3, rtol=1e-3, mask=upd),
            views_match=float((a["views"] == b["views"])[upd].mean()),
            rng_match=float((a["rand"] == b["rand"]).all(axis=-1)[upd].mean()),
            untouched_ok=float(np.all((a["planes"] == st["planes"]) | (np.isnan(a["planes"]) & np.isnan(st["planes"])), axis=-1)[keep].mean()),
        )
    return res, out, b


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_single_pass_photometric(model):
    scene = util.scene_of(model)
    H, W = scene.images[0].shape
    ctx, *_ = _mine(scene)
    ref = _ref(scene)
    ref.launch_init()
    res = {}
    r0, _, _ = _pass_compare(ctx, ref, 0, 0, H, W)
    res["black_it0"] = r0
    r1, _, _ = _pass_compare(ctx, ref, 1, 0, H, W)
    res["red_it0"] = r1
    ref.launch_pass(0, 1)
    r2, _, _ = _pass_compare(ctx, ref, 1, 1, H, W)
    res["red_it1"] = r2
    dump(f"pass_photo_{model}", res)
    for k, v in res.items():
        # as-compiled reading of the uninitialised plane variable (ACMMP.cu:1301) = the oracle's behaviour.  Measured (B200):
        # 0.9975 / 0.9932 / 0.9806 pinhole, 0.9975 / 0.9943 / 0.9901 sphere; red_it1 starts from a state one more racy
        # reference pass away
        assert v["mode1"]["plane_match"] >= (0.975 if k == "red_it1" else 0.988), (k, v)
        assert v["mode1"]["plane_match"] > v["mode0"]["plane_match"], (k, v)
        assert v["mode1"]["untouched_ok"] == 1.0, (k, v)
        assert v["mode1"]["rng_match"] >= 0.995, (k, v)


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_single_pass_geom(model):
    scene = util.scene_of(model)
    H, W = scene.images[0].shape
    imgs, cams, ids = scene.problem(0)
    rng = np.random.default_rng(2)
    depth_maps = [scene.depths_gt[i] * (1 + 0.01 * rng.standard_normal(scene.depths_gt[i].shape)).astype(np.float32) for i in ids]
    gt = util.random_planes(scene, 0, seed=1, perturb=0.02)
    prev = np.concatenate([util.world_normals(scene, 0, gt), depth_maps[0][..., None]], axis=-1).astype(np.float32)
    costs = rng.uniform(0.0, 0.6, (H, W)).astype(np.float32)
    ctx, *_ = _mine(scene)
    ctx.set_geom_consistency(False)
    ctx.set_depth_maps(depth_maps)
    ctx.set_planes(prev, costs)
    ref = _ref(scene, geom=True, depth_maps=depth_maps, prev_planes=prev, prev_costs=costs)
    # init (reload branch) must agree first
    ctx.random_init(); ctx.synchronize()
    ref.launch_init()
    a, b = ctx.download_state(), ref.download_state()
    res = dict(init=dict(planes=close_frac(a["planes"], b["planes"], 1e-5, 1e-5), costs=close_frac(a["costs"], b["costs"], 1e-4, 1e-4),
                         views=float((a["views"] == b["views"]).mean())))
    res["black_it0"], _, _ = _pass_compare(ctx, ref, 0, 0, H, W)
    res["red_it0"], _, _ = _pass_compare(ctx, ref, 1, 0, H, W)
    dump(f"pass_geom_{model}", res)
    assert res["init"]["planes"] >= 0.999 and res["init"]["costs"] >= 0.99, res
    for k in ("black_it0", "red_it0"):
        assert res[k]["mode1"]["plane_match"] >= 0.99, (k, res[k])             # measured 0.9954 .. 0.9965


def _race_sensitive(scene, st, colour, params, masks):
    """Pixels whose result of a planar-prior pass depends on the reference's data race (oracle.cpu_oracle.prior_pass_race:
    the CPU restatement run under both extreme interleavings)."""
    from oracle import cpu_oracle
    imgs, cams, _ = scene.problem(0)
    H, W = masks.shape
    m = masks.astype(np.uint32)
    pp = np.zeros((H, W, 4), np.float32)
    pp[m > 0] = np.asarray(params, np.float32)[m[m > 0] - 1]
    _, _, sens = cpu_oracle.prior_pass_race(imgs, cams, st, colour, 0, pp, m, hierarchy=True)
    return sens


def test_single_pass_prior_and_hierarchy():
    """Planar-prior stage on top of a hierarchy stage (the level > 0 schedule, main.cpp:453-458).
    Hierarchy passes must agree like photometric ones.  Planar-prior passes cannot: there the reference has a data race
    that FIRES -- a thread stores plane_hypotheses[center] in the middle of its work (ACMMP.cu:1283, :1295) while threads
    of other warps re-read that same-colour pixel at the same point of theirs (:1262, :1279, :1291); without the prior the
    re-read comes long before anybody's store and the race is benign (tests/test_cpu_oracle.py::
    test_prior_pass_gap_is_the_references_data_race has the evidence on the reference's own vectors).  This library reads
    the pre-pass state throughout (double-buffered), one legal outcome.  Asserted: on the pixels whose result does not
    depend on the race the agreement is that of a photometric pass; overall it is the measured 95 / 92 % minus a margin;
    and the reference does not even reproduce itself across two launches from the same state on the sensitive pixels."""
    scene = util.pinhole_scene()
    H, W = scene.images[0].shape
    rng = np.random.default_rng(3)
    gt = util.random_planes(scene, 0, seed=1, perturb=0.03)
    nw = util.world_normals(scene, 0, gt)
    coarse_normals = np.ascontiguousarray(nw[::2, ::2][: H // 2, : W // 2])
    coarse_costs = rng.uniform(0.0, 0.5, (H // 2, W // 2)).astype(np.float32)
    fine_depth = scene.depths_gt[0] * (1 + 0.02 * rng.standard_normal((H, W))).astype(np.float32)
    coarse4 = np.concatenate([coarse_normals, coarse_costs[..., None]], axis=-1)

    ctx, *_ = _mine(scene)
    ctx.set_hierarchy()
    ctx.set_hierarchy_inputs(coarse4, fine_depth)
    ref = _ref(scene, hierarchy=True, coarse_normals=coarse_normals, coarse_costs=coarse_costs, fine_depth=fine_depth)
    ctx.random_init(); ctx.synchronize()
    ref.launch_init()
    a, b = ctx.download_state(), ref.download_state(pre_costs=True)
    res = dict(init_upsample=dict(planes=close_frac(a["planes"], b["planes"], 1e-4, 1e-4),
                                  costs=close_frac(a["costs"], b["costs"], 1e-4, 1e-4),
                                  pre_costs=close_frac(a["pre_costs"], b["pre_costs"], 1e-4, 1e-4),
                                  views=float((a["views"] == b["views"]).mean())))
    res["hier_black"], _, _ = _pass_compare(ctx, ref, 0, 0, H, W)
    res["hier_red"], _, _ = _pass_compare(ctx, ref, 1, 0, H, W)
    # finish the stage on the reference, then the prior stage on the same objects
    ref.launch_finalize()
    st = ref.download_state(pre_costs=True)
    params, masks = util.grid_prior(scene, 0)
    ref.set_prior(params, masks)
    ctx.upload_state(planes=st["planes"], costs=st["costs"], views=st["views"], rand=st["rand"], pre_costs=st["pre_costs"])
    ctx.set_planar_prior_inputs(params, masks)
    ctx.random_init(); ctx.synchronize()
    ref.launch_init()
    a, b = ctx.download_state(), ref.download_state()
    res["init_prior"] = dict(planes=close_frac(a["planes"], b["planes"], 1e-4, 1e-4),
                             costs=close_frac(a["costs"], b["costs"], 1e-4, 1e-4),
                             rng=float((a["rand"] == b["rand"]).all(axis=-1).mean()))
    for name, colour in (("prior_black", 0), ("prior_red", 1)):
        st0 = ref.download_state(rand=True, pre_costs=True)
        sens = _race_sensitive(scene, st0, colour, params, masks)
        r, out, b1 = _pass_compare(ctx, ref, colour, 0, H, W)
        # the reference against ITSELF: the same launch again from the same state
        ref.upload_state(planes=st0["planes"], costs=st0["costs"], views=st0["views"], rand=st0["rand"], pre_costs=st0["pre_costs"])
        ref.launch_pass(colour, 0)
        b2 = ref.download_state()
        upd = util.colour_mask(H, W, colour) & util.interior(H, W, 4)

        def same(p, q):
            return np.all((np.abs(p - q) <= 1e-4 + 1e-4 * np.abs(q)) | (np.isnan(p) & np.isnan(q)), axis=-1)
        mine_ok = same(out[1]["planes"], b1["planes"])
        self_ok = same(b2["planes"], b1["planes"])
        r["race"] = dict(sensitive_frac=float(sens[upd].mean()),
                         mine_on_insensitive=float(mine_ok[upd & ~sens].mean()), mine_on_sensitive=float(mine_ok[upd & sens].mean()),
                         ref_vs_ref=float(self_ok[upd].mean()), ref_vs_ref_on_insensitive=float(self_ok[upd & ~sens].mean()),
                         ref_vs_ref_on_sensitive=float(self_ok[upd & sens].mean()))
        res[name] = r
    dump("pass_prior_hier", res)
    assert res["init_upsample"]["planes"] >= 0.999 and res["init_upsample"]["pre_costs"] >= 0.99, res
    assert res["init_prior"]["planes"] >= 0.999 and res["init_prior"]["rng"] == 1.0, res
    for k in ("hier_black", "hier_red"):
        assert res[k]["mode1"]["plane_match"] >= 0.985, (k, res[k])            # measured 0.992 / 0.991
    # measured 0.953 / 0.920 overall (round 1), minus 0.5 %
    assert res["prior_black"]["mode1"]["plane_match"] >= 0.948, res["prior_black"]
    assert res["prior_red"]["mode1"]["plane_match"] >= 0.915, res["prior_red"]
    for k in ("prior_black", "prior_red"):
        assert res[k]["race"]["mine_on_insensitive"] >= 0.98, (k, res[k])
        assert res[k]["race"]["mine_on_sensitive"] < res[k]["race"]["mine_on_insensitive"], (k, res[k])


# ------------------------------------------------------------------------------------------
# GetDepthandNormal + median filter
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_finalize_matches_reference(model):
    scene = util.scene_of(model)
    ctx, *_ = _mine(scene)
    ref = _ref(scene)
    ref.launch_init()
    ref.launch_pass(0, 0)
    st = ref.download_state()
    ctx.upload_state(planes=st["planes"], costs=st["costs"], views=st["views"], rand=st["rand"])
    ctx.finalize(); ctx.synchronize()
    ref.launch_finalize()
    a, b = ctx.download_state(), ref.download_state()
    res = dict(depth=close_frac(a["planes"][..., 3], b["planes"][..., 3], 0, 1e-5),
               depth_1e3=close_frac(a["planes"][..., 3], b["planes"][..., 3], 0, 1e-3),
               normal=close_frac(a["planes"][..., :3], b["planes"][..., :3], 1e-6, 1e-5))
    dump(f"finalize_{model}", res)
    # planes nearly parallel to the ray have an ill-conditioned depth: a few pixels in 1e4 exceed 1e-5
    assert res["depth"] >= 0.999 and res["depth_1e3"] >= 0.9999 and res["normal"] >= 0.9999, res


# ------------------------------------------------------------------------------------------
# whole stage, statistical
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_full_stage_statistical(model):
    scene = util.scene_of(model)
    H, W = scene.images[0].shape
    ctx, *_ = _mine(scene)
    ref = _ref(scene)
    ctx.run_patch_match()
    pa, ca = ctx.get_result()
    ref_ms = ref.run_patch_match()
    pb, cb = ref.get_result()
    m = util.interior(H, W, 8)
    gt = scene.depths_gt[0]
    rel = np.abs(pa[..., 3] - pb[..., 3]) / np.maximum(np.abs(pb[..., 3]), 1e-9)
    ang = util.angle_deg(pa[..., :3], pb[..., :3])
    res = dict(
        depth_within_1pct=float((rel <= 0.01)[m].mean()), normal_within_5deg=float((ang <= 5.0)[m].mean()),
        mine_vs_gt_1pct=float((np.abs(pa[..., 3] - gt) / gt <= 0.01)[m].mean()),
        ref_vs_gt_1pct=float((np.abs(pb[..., 3] - gt) / gt <= 0.01)[m].mean()),
        mean_cost_mine=float(np.nanmean(ca[m])), mean_cost_ref=float(np.nanmean(cb[m])),
        ref_ms=ref_ms, mine=ctx.timings(),
    )
    dump(f"full_stage_{model}", res)
    # both converge to the same surface; the quality of mine must not be below the reference's
    assert res["mine_vs_gt_1pct"] >= res["ref_vs_gt_1pct"] - 0.02, res
    # measured (B200): pinhole 99.96 % of the pixels within 1 % depth / 98.3 % of the normals within 5 deg, sphere 99.6 % / 95.6 %;
    # the reference against itself on a single photometric stage: tests/test_gpu_pipeline_parity.py (C1)
    assert res["depth_within_1pct"] >= 0.99, res
    assert res["normal_within_5deg"] >= (0.97 if model == "pinhole" else 0.94), res
