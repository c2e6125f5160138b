"""Golden vectors of the fusion stage: small seeded scenes through the REFERENCE's SimpleFusionKernel (unmodified, compiled
into oracle/_ref/libacmmp_ref.so, behind RunFusionCuda's texture set-up in oracle/ref_harness.cu) on a GPU box.
tests/test_cpu_oracle.py pins oracle/acmmp_oracle.c: orc_fuse_view against them.

    python tests/golden/make_fusion_golden.py [outdir]        # default: gpurun_out/golden"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "acmmp-spherical_b200"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def inputs(model):
    """Deterministic inputs shared with the test: depth = ground truth with noise and holes, normals = ground-truth planes'
    normals with noise, colour images with three different channels."""
    import util
    from acmmp_b200 import synth
    scene = (synth.make_pinhole_scene(n_views=4, width=128, height=96, focal=100.0, seed=11) if model == "pinhole"
             else synth.make_sphere_scene(n_views=4, width=160, height=80, seed=12))
    rng = np.random.default_rng(21)
    depths, normals, colours = [], [], []
    for v in range(len(scene.images)):
        d = scene.depths_gt[v] * (1.0 + 0.002 * rng.standard_normal(scene.depths_gt[v].shape)).astype(np.float32)
        d[rng.random(d.shape) < 0.02] = 0.0
        n = util.world_normals(scene, v, synth.gt_planes(scene, v))
        n += 0.02 * rng.standard_normal(n.shape).astype(np.float32)
        n /= np.linalg.norm(n, axis=-1, keepdims=True)
        g = np.clip(scene.images[v], 0, 255).astype(np.uint8)
        depths.append(np.ascontiguousarray(d, np.float32))
        normals.append(np.ascontiguousarray(n, np.float32))
        colours.append(np.ascontiguousarray(np.stack([g, np.roll(g, 5, axis=1), 255 - g], axis=-1)))
    return scene, depths, normals, colours


def main():
    from oracle.ref_driver import RefFusion
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    for model in ("pinhole", "sphere"):
        scene, depths, normals, colours = inputs(model)
        store = {}
        for tag, col in (("grey", None), ("colour", colours)):
            ref = RefFusion(scene.cams, depths, normals, scene.images, colours=col)
            for r in (0, 2):
                h, w = depths[r].shape
                pts, flags = ref.run(r, list(scene.pairs[r][1]))
                dense = np.zeros((h * w, 9), np.float32)
                dense[flags.ravel() != 0] = pts
                store[f"{tag}_view{r}_points"] = dense.reshape(h, w, 9)
                store[f"{tag}_view{r}_flags"] = flags.astype(np.uint8)
            ref.close()
        np.savez_compressed(out / f"golden_fusion_{model}.npz", **store)
        print(model, {k: v.shape for k, v in store.items()}, "points", int(store["grey_view0_flags"].sum()))


if __name__ == "__main__":
    main()
