"""CPU tests: the plain-C restatement (oracle/acmmp_oracle.c) pinned against outputs of the REFERENCE
ITSELF (tests/golden/*.npz, produced on a B200 by tests/golden/make_golden.py from the unmodified
reference sources compiled for sm_100).  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so these vectors are the pin.

Tolerances: the reference is built with --use_fast_math and reads images through the texture unit
(1.8 fixed-point bilinear fractions); libm + an emulated filter agree to ~1e-4, not bit-exactly.
Integer work (XORWOW states, view bit masks) must match exactly.
"""
import ctypes as C
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
