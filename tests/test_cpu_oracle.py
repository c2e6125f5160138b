"""CPU tests: the plain-C restatement (oracle/acmmp_oracle.c) pinned against outputs of the REFERENCE
ITSELF (tests/golden/*.npz, produced on a B200 by tests/golden/make_golden.py from the unmodified
reference sources compiled for sm_100).  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so these vectors are the pin.

Tolerances: the reference is built with --use_fast_math and reads images through the texture unit
(1.8 fixed-point bilinear fractions); libm + an emulated filter agree to ~1e-4, not bit-exactly.
Integer work (XORWOW states, view bit masks) must match exactly.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

import util
from util import close_frac

GOLD = Path(__file__).resolve().parent / "golden"


def load(model):
    from acmmp_b200 import Camera
    g = np.load(GOLD / f"golden_{model}.npz")
    imgs = [g["images"][i].astype(np.float32) for i in range(g["images"].shape[0])]
    cams = [Camera.from_buffer_copy(g["cams"][i].tobytes()) for i in range(g["cams"].shape[0])]
    return g, imgs, cams


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_warp_and_ncc_against_reference_vectors(model):
    from oracle import cpu_oracle as co
    g, imgs, cams = load(model)
    planes = g["probe_planes"]
    for v in (1, 2, 3):
        w = co.warp_map(cams, planes, v)
        assert close_frac(w[..., :2], g[f"warp_v{v}"][..., :2], atol=2e-3, rtol=1e-5) >= 0.999
        assert close_frac(w[..., 2:], g[f"warp_v{v}"][..., 2:], atol=0, rtol=1e-4) >= 0.999
        c = co.ncc_map(imgs, cams, planes, v)
        ref = g[f"ncc_v{v}"]
        # measured: PINHOLE 99.99 % within 1e-3 / ~90 % within 1e-4; SPHERE 99 % / ~48 % (its angular
        # bilateral weights are exp(-x) with x ~ 10..100, so MUFU sin/cos/ex2 vs libm shows up earlier)
        assert close_frac(c, ref, atol=1e-3, rtol=1e-3) >= (0.999 if model == "pinhole" else 0.98), (model, v, float(np.abs(c - ref).max()))
        assert close_frac(c, ref, atol=1e-4, rtol=1e-4) >= (0.85 if model == "pinhole" else 0.4)
        assert abs(float((c >= 2).mean()) - float((ref >= 2).mean())) < 2e-3


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_initial_cost_and_view_masks(model):
    from oracle import cpu_oracle as co
    g, imgs, cams = load(model)
    c, views = co.initcost_map(imgs, cams, g["probe_planes"])
    assert close_frac(c, g["initcost"], atol=1e-3, rtol=1e-3) >= 0.995
    assert float((views == g["initcost_views"]).mean()) >= 0.99


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_geometric_consistency_cost(model):
    from oracle import cpu_oracle as co
    g, imgs, cams = load(model)
    dms = [g["geom_depth_maps"][i] for i in range(g["geom_depth_maps"].shape[0])]
    for v in (1, 3):
        c = co.geom_map(dms, cams, g["probe_planes"], v)
        assert close_frac(c, g[f"geomcost_v{v}"], atol=5e-3, rtol=1e-3) >= 0.99


def test_xorwow_states_are_bit_exact():
    """curand_init(seed, y, x) + the draws of RandomInitialization: integer arithmetic, exact."""
    from oracle import cpu_oracle as co
    g, imgs, cams = load("pinhole")
    st = co.random_init(imgs, cams, int(g["seed"]), with_costs=False)
    assert np.array_equal(st["rand"], g["init_rand"])


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_random_init_planes_and_costs(model):
    from oracle import cpu_oracle as co
    g, imgs, cams = load(model)
    st = co.random_init(imgs, cams, int(g["seed"]), with_costs=True)
    assert np.array_equal(st["rand"], g["init_rand"])
    assert close_frac(st["planes"], g["init_planes"], atol=2e-5, rtol=2e-5) >= 0.999
    assert close_frac(st["costs"], g["init_costs"], atol=1e-3, rtol=1e-3) >= 0.99
    assert float((st["views"] == g["init_views"]).mean()) >= 0.99


def test_jbu_against_reference_vector():
    from oracle import cpu_oracle as co
    g, imgs, cams = load("pinhole")
    out = co.jbu(imgs[0], g["jbu_coarse"])
    assert close_frac(out, g["jbu_out"], atol=0, rtol=1e-5) >= 0.9999


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_depth_normal_and_median_filter(model):
    from oracle import cpu_oracle as co
    g, imgs, cams = load(model)
    pl = co.depth_normal(cams[0], g["red0_planes"])
    pl = co.median_filter(pl, g["red0_costs"], 0)
    pl = co.median_filter(pl, g["red0_costs"], 1)
    ref = g["final_planes"]
    m = util.interior(*ref.shape[:2], 4)
    assert close_frac(pl[..., 3], ref[..., 3], atol=0, rtol=1e-4, mask=m) >= 0.999
    assert close_frac(pl[..., :3], ref[..., :3], atol=1e-5, rtol=1e-4, mask=np.repeat(m[..., None], 3, -1)) >= 0.999


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_checkerboard_pass_against_reference_state(model):
    """One black pass from the reference's own post-initialisation state.  Pixel-exact agreement is not
    possible (fast-math vs libm flips arg-min decisions, the reference races on same-colour reads), so the
    test asserts the agreement rate, with the as-compiled reading of the uninitialised plane variable
    (ACMMP.cu:1301) clearly ahead of the intended one."""
    from oracle import cpu_oracle as co
    g, imgs, cams = load(model)
    st = dict(planes=g["init_planes"], costs=g["init_costs"], views=g["init_views"], rand=g["init_rand"], pre_costs=None)
    H, W = st["costs"].shape
    upd = util.colour_mask(H, W, 0) & util.interior(H, W, 4)
    rate = {}
    for mode in (True, False):
        out = co.checkerboard_pass(imgs, cams, st, 0, 0, as_compiled=mode)
        same = np.all(np.abs(out["planes"] - g["black0_planes"]) <= 1e-3 + 1e-3 * np.abs(g["black0_planes"]), axis=-1)
        rate[mode] = float(same[upd].mean())
        if mode:
            assert float((out["rand"] == g["black0_rand"]).all(-1)[upd].mean()) >= 0.97
            assert float((out["views"] == g["black0_views"])[upd].mean()) >= 0.97
    assert rate[True] >= 0.93, rate
    assert rate[True] > rate[False], rate


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_prior_pass_gap_is_the_references_data_race(model):
    """Round-1 finding: a planar-prior pass from an identical state matched the reference for only 92-95 % of the planes
    (photometric / geometric / hierarchy passes: 98-99.7 %).  Root cause: in prior mode the reference writes
    plane_hypotheses[center] in the middle of a thread's work (ACMMP.cu:1283, :1295) and other threads re-read
    same-colour neighbours at the same point of theirs (:1262, :1279, :1291) -- a data race whose outcome depends on warp
    timing (cpu_oracle.prior_pass_race).  Evidence on the reference's own prior-pass vectors: on the pixels whose result
    does not depend on the race the restatement agrees with the reference as well as it does in a photometric pass;
    the race-sensitive pixels follow the `early` reading for ~75 % and the `late` one for most of the rest."""
    from oracle import cpu_oracle as co
    g, imgs, cams = load(model)
    st = dict(planes=g["prior_init_planes"], costs=g["prior_init_costs"], views=g["prior_init_views"], rand=g["prior_init_rand"],
              pre_costs=g["prior_in_pre_costs"])
    H, W = st["costs"].shape
    upd = util.colour_mask(H, W, 0) & util.interior(H, W, 4)
    masks = g["prior_masks"].astype(np.uint32)
    pp = np.zeros((H, W, 4), np.float32)
    pp[masks > 0] = g["prior_params"][masks[masks > 0] - 1]
    early, late, sens = co.prior_pass_race(imgs, cams, st, 0, 0, pp, masks)
    ref = g["prior_black0_planes"]

    def same(a):
        return np.all(np.abs(a - ref) <= 1e-3 + 1e-3 * np.abs(ref), axis=-1)
    e, l = same(early["planes"]), same(late["planes"])
    # the same restatement on the photometric pass of the same scene: its numerical floor (libm vs fast-math)
    st0 = dict(planes=g["init_planes"], costs=g["init_costs"], views=g["init_views"], rand=g["init_rand"], pre_costs=None)
    photo = co.checkerboard_pass(imgs, cams, st0, 0, 0)
    floor = float(np.all(np.abs(photo["planes"] - g["black0_planes"]) <= 1e-3 + 1e-3 * np.abs(g["black0_planes"]), axis=-1)[upd].mean())
    res = dict(early=float(e[upd].mean()), late=float(l[upd].mean()), either=float((e | l)[upd].mean()),
               sensitive_frac=float(sens[upd].mean()), early_on_insensitive=float(e[upd & ~sens].mean()),
               early_on_sensitive=float(e[upd & sens].mean()), late_on_sensitive=float(l[upd & sens].mean()), photometric_floor=floor)
    util.dump(f"oracle_prior_race_{model}", res)
    assert res["early_on_insensitive"] >= floor - 0.03, res          # measured: 0.993 vs 0.991 (pinhole), 0.942 vs 0.965 (sphere)
    assert res["either"] >= res["early"] + 0.04, res                 # measured: 0.970 vs 0.921, 0.920 vs 0.845
    assert res["early_on_sensitive"] < res["early_on_insensitive"] - 0.1, res


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_fusion_restatement_against_reference_vectors(model):
    """orc_fuse_view (SimpleFusionKernel, ACMMP.cu:1664-1814, restated in C) against what the compiled reference kernel
    produced on a B200 for the same inputs (tests/golden/make_fusion_golden.py): the same pixels yield a point -- apart from
    pixels that sit on one of the three consistency thresholds, where libm's acos / hypot and the device's differ in the
    last bits --, coordinates, normals and colours agree to 1e-4, with grey and with three-channel colour images."""
    sys.path.insert(0, str(GOLD))
    from make_fusion_golden import inputs
    from oracle import cpu_oracle
    scene, depths, normals, colours = inputs(model)
    gold = np.load(GOLD / f"golden_fusion_{model}.npz")
    for tag, col in (("grey", None), ("colour", colours)):
        for r in (0, 2):
            pts, flags = cpu_oracle.fuse_view(scene.cams, depths, normals, scene.images, r, list(scene.pairs[r][1]), colours=col)
            want_pts, want_flags = gold[f"{tag}_view{r}_points"], gold[f"{tag}_view{r}_flags"].astype(bool)
            assert want_flags.sum() > 1000
            assert (flags == want_flags).mean() >= 0.9995, (tag, r, (flags == want_flags).mean())
            both = flags & want_flags
            scale = np.abs(want_pts[both][:, :3]).max()
            # (a pixel where ONE source view sits on a threshold keeps its point but averages over one view more or less:
            # a handful per ten thousand)
            assert (np.abs(pts[both][:, :3] - want_pts[both][:, :3]) <= 1e-4 * scale).all(axis=1).mean() >= 0.9995
            assert (np.abs(pts[both][:, 3:6] - want_pts[both][:, 3:6]) <= 1e-4).all(axis=1).mean() >= 0.9995
            assert (np.abs(pts[both][:, 6:] - want_pts[both][:, 6:]) <= 0.05).all(axis=1).mean() >= 0.9995
            if col is not None:                                   # three different channels, in the kernel's B, G, R order
                assert np.abs(want_pts[both][:, 6] - want_pts[both][:, 8]).mean() > 10
