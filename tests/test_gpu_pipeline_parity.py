"""End-to-end parity of the WHOLE multi-scale schedule against the unmodified reference (north_star's full-map bar:
>= 99 % of pixels within 1 % relative depth, normals within 5 degrees).

Both arms run the reference's schedule for one reference view (main.cpp:417-476): per pyramid level photometric stage
(hierarchy + JBU hand-over above the coarsest level) -> CPU planar prior -> prior stage -> two geometric stages, each
arm feeding ITS OWN results forward (its own prior from its own photometric result, its own JBU):
  * product : acmmp_b200.pipeline.B200Backend  (libacmmp_b200.so, GPU-resident between stages)
  * oracle  : oracle.ref_pipeline.ReferenceBackend (oracle/_ref/libacmmp_ref.so, the reference's ACMMP.cu + ACMMP.cpp
              compiled for sm_100, one ACMMP object per stage, .dmb files in between, RunJBU)
The neighbours' depth maps of the geometric stages are the same rendered maps on both sides.

PatchMatch is a randomised search and the reference is not reproducible against itself: cuRAND is seeded from clock64()
(ACMMP.cu:684), the near candidates race on same-colour pixels (:1047-1140), the planar-prior pass races on planes
(tests/test_gpu_parity.py::test_single_pass_prior_and_hierarchy).  So every comparison is reported next to the
reference's own NOISE FLOOR on the same scene -- the reference run twice with different seeds -- and asserted as
min(north_star bar, floor - margin): the product must agree with the reference at least as well as the reference
agrees with itself.
"""
import numpy as np
import pytest

import util
from util import dump

pytestmark = pytest.mark.gpu


def _agreement(pa, pb, border=8, valid=None):
    H, W = pa.shape[:2]
    m = util.interior(H, W, border)
    if valid is not None:
        m &= valid
    rel = np.abs(pa[..., 3] - pb[..., 3]) / np.maximum(np.abs(pb[..., 3]), 1e-9)
    ang = util.angle_deg(pa[..., :3], pb[..., :3])
    return dict(depth_within_1pct=float((rel <= 0.01)[m].mean()), normal_within_5deg=float((ang <= 5.0)[m].mean()),
                both=float(((rel <= 0.01) & (ang <= 5.0))[m].mean()))


def _vs_gt(p, gt, border=8):
    H, W = gt.shape
    m = util.interior(H, W, border)
    return float((np.abs(p[..., 3] - gt) / gt <= 0.01)[m].mean())


def _run(levels, backend):
    from acmmp_b200.pipeline import run_view
    planes, costs = run_view(levels, backend, prior_cache={})
    planes, costs = np.array(planes), np.array(costs)
    backend.end()
    return planes, costs


def _scene(model):
    from acmmp_b200 import synth
    if model == "pinhole":      # 1280 x 960 -> levels 640 x 480 and 1280 x 960
        return synth.make_pinhole_scene(n_views=5, width=1280, height=960, focal=1000.0, seed=11)
    return synth.make_sphere_scene(n_views=5, width=2048, height=1024, seed=12)      # levels 512 x 256, 1024 x 512, 2048 x 1024


@pytest.mark.parametrize("model", ["pinhole", "sphere"])
def test_multiscale_schedule_agrees_with_the_reference_at_the_north_star_tolerance(model):
    from acmmp_b200 import pipeline
    from oracle.ref_pipeline import ReferenceBackend
    scene = _scene(model)
    levels = pipeline.build_levels(scene, 0)
    assert len(levels) >= 2
    gt = scene.depths_gt[0]
    mine, mine_c = _run(levels, pipeline.B200Backend(0, seed=1234))
    ref_a, ref_ac = _run(levels, ReferenceBackend(0, seed=1234))
    ref_b, _ = _run(levels, ReferenceBackend(0, seed=4321))                 # the reference's own noise floor
    mine_b, _ = _run(levels, pipeline.B200Backend(0, seed=4321))
    res = dict(levels=[list(l.images[0].shape[::-1]) for l in levels],
               mine_vs_ref=_agreement(mine, ref_a), ref_vs_ref_other_seed=_agreement(ref_b, ref_a),
               mine_vs_mine_other_seed=_agreement(mine_b, mine), mine_other_seed_vs_ref=_agreement(mine_b, ref_a),
               mine_vs_gt_1pct=_vs_gt(mine, gt), ref_vs_gt_1pct=_vs_gt(ref_a, gt),
               mean_cost_mine=float(np.nanmean(mine_c)), mean_cost_ref=float(np.nanmean(ref_ac)))
    # pixels where the reference agrees with itself across seeds = where the scene determines the answer
    rel = np.abs(ref_b[..., 3] - ref_a[..., 3]) / np.maximum(np.abs(ref_a[..., 3]), 1e-9)
    stable = (rel <= 0.01) & (util.angle_deg(ref_b[..., :3], ref_a[..., :3]) <= 5.0)
    res["mine_vs_ref_where_ref_is_reproducible"] = _agreement(mine, ref_a, valid=stable)
    dump(f"pipeline_parity_{model}", res)
    floor = res["ref_vs_ref_other_seed"]
    got = res["mine_vs_ref"]
    for key in ("depth_within_1pct", "normal_within_5deg"):
        assert got[key] >= min(0.99, floor[key] - 0.01), (key, res)
    # where the reference reproduces itself, the north_star bar holds outright
    stable_got = res["mine_vs_ref_where_ref_is_reproducible"]
    assert stable_got["depth_within_1pct"] >= 0.99 and stable_got["normal_within_5deg"] >= 0.99, res
    assert res["mine_vs_gt_1pct"] >= res["ref_vs_gt_1pct"] - 0.01, res


def test_c1_single_scale_acmh_equivalent_640x480_5_views():
    """BASELINE.json configs[0]: 5 views 640 x 480, single-scale, no geometric consistency / prior = the state after the
    first RunPatchMatch of ProcessProblem (main.cpp:94-111; not reachable through ./ACMMP)."""
    from acmmp_b200 import synth, Context
    from oracle.ref_driver import RefACMMP
    scene = synth.make_pinhole_scene(n_views=5, width=640, height=480, focal=500.0, seed=1)
    imgs, cams, ids = scene.problem(0)
    out = {}
    for seed in (1234, 4321):
        ctx = Context(0)
        ctx.set_views(imgs, cams)
        ctx.set_seed(seed)
        ctx.run_patch_match()
        out["mine", seed] = tuple(np.array(a) for a in ctx.get_result())
        ctx.close()
        ref = RefACMMP(imgs, cams, seed=seed)
        ref.run_patch_match()
        out["ref", seed] = ref.get_result()
        ref.close()
    gt = scene.depths_gt[0]
    res = dict(mine_vs_ref=_agreement(out["mine", 1234][0], out["ref", 1234][0]),
               ref_vs_ref_other_seed=_agreement(out["ref", 4321][0], out["ref", 1234][0]),
               mine_vs_gt_1pct=_vs_gt(out["mine", 1234][0], gt), ref_vs_gt_1pct=_vs_gt(out["ref", 1234][0], gt))
    dump("pipeline_parity_c1", res)
    floor, got = res["ref_vs_ref_other_seed"], res["mine_vs_ref"]
    for key in ("depth_within_1pct", "normal_within_5deg"):
        assert got[key] >= min(0.99, floor[key] - 0.01), (key, res)
    assert res["mine_vs_gt_1pct"] >= res["ref_vs_gt_1pct"] - 0.01, res
