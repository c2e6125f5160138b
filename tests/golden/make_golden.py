"""Generate the golden vectors under tests/golden/ by running the REFERENCE ITSELF.

Runs on a GPU box (needs oracle/_ref/libacmmp_ref.so = the unmodified reference sources compiled
for sm_100, see oracle/Makefile): small seeded synthetic scenes go through the reference's own
kernels and host set-up code, and inputs + outputs are stored as compressed .npz files.  The CPU
tests (tests/test_cpu_oracle.py) pin oracle/acmmp_oracle.c against these vectors; the GPU tests
compare the CUDA path with them as well.

    python tests/golden/make_golden.py [outdir]        # default: gpurun_out/golden
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "acmmp-spherical_b200"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SEED = 4321


def cam_array(cams):
    return np.stack([np.frombuffer(bytes(c), dtype=np.uint8) for c in cams])


def scenario(model):
    from acmmp_b200 import synth
    if model == "pinhole":
        return synth.make_pinhole_scene(n_views=4, width=128, height=96, focal=100.0, seed=11)
    return synth.make_sphere_scene(n_views=4, width=160, height=80, seed=12)


def main():
    import cv2
    import util
    from oracle.ref_driver import RefACMMP, run_jbu
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    for model in ("pinhole", "sphere"):
        scene = scenario(model)
        imgs, cams, ids = scene.problem(0)
        H, W = imgs[0].shape
        g = dict(images=np.stack(imgs).astype(np.uint8), cams=cam_array(cams), seed=np.uint64(SEED))
        rng = np.random.default_rng(5)

        # ---- deterministic sub-kernels at fixed planes
        ref = RefACMMP(imgs, cams, seed=SEED)
        planes = util.random_planes(scene, 0, seed=5, perturb=0.05)
        g["probe_planes"] = planes
        for v in (1, 2, 3):
            g[f"ncc_v{v}"] = ref.probe_ncc(planes, v)
            g[f"warp_v{v}"] = ref.probe_warp(planes, v)
        c, sv = ref.probe_initcost(planes)
        g["initcost"], g["initcost_views"] = c, sv

        # ---- RandomInitialization + one black pass + one red pass + finalize (photometric stage)
        ref.launch_init()
        st = ref.download_state()
        for k in ("planes", "costs", "views", "rand"):
            g[f"init_{k}"] = st[k]
        ref.launch_pass(0, 0)
        st = ref.download_state()
        for k in ("planes", "costs", "views", "rand"):
            g[f"black0_{k}"] = st[k]
        ref.launch_pass(1, 0)
        st = ref.download_state()
        for k in ("planes", "costs", "views", "rand"):
            g[f"red0_{k}"] = st[k]
        ref.launch_finalize()
        st = ref.download_state()
        g["final_planes"], g["final_costs"] = st["planes"], st["costs"]
        ref.close()

        # ---- geometric consistency: probe + one pass
        depth_maps = [scene.depths_gt[i] * (1 + 0.01 * rng.standard_normal(scene.depths_gt[i].shape)).astype(np.float32) for i in ids]
        gt = util.random_planes(scene, 0, seed=1, perturb=0.02)
        prev = np.concatenate([util.world_normals(scene, 0, gt), depth_maps[0][..., None]], axis=-1).astype(np.float32)
        prev_costs = rng.uniform(0.0, 0.6, (H, W)).astype(np.float32)
        g["geom_depth_maps"] = np.stack(depth_maps)
        g["geom_prev_planes"], g["geom_prev_costs"] = prev, prev_costs
        ref = RefACMMP(imgs, cams, seed=SEED, geom=True, depth_maps=depth_maps, prev_planes=prev, prev_costs=prev_costs)
        for v in (1, 3):
            g[f"geomcost_v{v}"] = ref.probe_geom(planes, v)
        ref.launch_init()
        st = ref.download_state()
        for k in ("planes", "costs", "views", "rand"):
            g[f"geom_init_{k}"] = st[k]
        ref.launch_pass(0, 0)
        st = ref.download_state()
        for k in ("planes", "costs", "views", "rand"):
            g[f"geom_black0_{k}"] = st[k]
        ref.close()

        # ---- hierarchy (upsample) stage followed by the planar-prior stage on the same object
        nw = util.world_normals(scene, 0, util.random_planes(scene, 0, seed=1, perturb=0.03))
        coarse_normals = np.ascontiguousarray(nw[::2, ::2][: H // 2, : W // 2])
        coarse_costs = rng.uniform(0.0, 0.5, (H // 2, W // 2)).astype(np.float32)
        fine_depth = scene.depths_gt[0] * (1 + 0.02 * rng.standard_normal((H, W))).astype(np.float32)
        g["hier_coarse_normals"], g["hier_coarse_costs"], g["hier_fine_depth"] = coarse_normals, coarse_costs, fine_depth
        ref = RefACMMP(imgs, cams, seed=SEED, hierarchy=True, coarse_normals=coarse_normals, coarse_costs=coarse_costs, fine_depth=fine_depth)
        ref.launch_init()
        st = ref.download_state(pre_costs=True)
        for k in ("planes", "costs", "views", "rand", "pre_costs"):
            g[f"hier_init_{k}"] = st[k]
        ref.launch_pass(0, 0)
        st = ref.download_state()
        for k in ("planes", "costs", "views", "rand"):
            g[f"hier_black0_{k}"] = st[k]
        ref.launch_pass(1, 0)
        ref.launch_finalize()
        params, masks = util.grid_prior(scene, 0, cell=8)
        g["prior_params"], g["prior_masks"] = params, masks
        ref.set_prior(params, masks)
        st = ref.download_state(pre_costs=True)
        for k in ("planes", "costs", "views", "rand", "pre_costs"):
            g[f"prior_in_{k}"] = st[k]
        ref.launch_init()
        st = ref.download_state()
        for k in ("planes", "costs", "views", "rand"):
            g[f"prior_init_{k}"] = st[k]
        ref.launch_pass(0, 0)
        st = ref.download_state()
        for k in ("planes", "costs", "views", "rand"):
            g[f"prior_black0_{k}"] = st[k]
        ref.close()

        # ---- JBU
        coarse = cv2.resize(scene.depths_gt[0], (W // 2, H // 2), interpolation=cv2.INTER_NEAREST)
        g["jbu_coarse"] = coarse
        g["jbu_out"] = run_jbu(imgs[0], coarse)

        np.savez_compressed(out / f"golden_{model}.npz", **g)
        print("wrote", out / f"golden_{model}.npz", {k: v.shape for k, v in g.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
