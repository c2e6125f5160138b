"""Per-view multi-scale pipeline driver over the C ABI.

Follows the reference's schedule for ONE reference view (main.cpp:417-476 + ProcessProblem,
main.cpp:73-210): per pyramid level, [JBU of the previous level's depth] -> photometric stage
(hierarchy at levels > 0) -> CPU planar prior -> planar-prior stage on the same object ->
geometric-consistency stage x 2.  The reference passes state between stages through .dmb files and
builds a new ACMMP object per stage; here one context per level is re-used and state is handed
over as arrays.  The C++ twin of this file is ../host/acmmp_driver.cpp.

Neighbour depth maps for the geometric stages come from `neighbour_depths` (the per-view results of
the other views in a full run; benchmark runs use rendered stand-ins).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from . import Context, synth
from .prior import planar_prior


@dataclass
class Level:
    images: list            # [ref, src...] float32
    cams: list
    neighbour_depths: list  # depth maps of the source views at this level (float32), len == n-1


@dataclass
class StageTimes:
    gpu_ms: float = 0.0          # CUDA-event time of the kernels (inputs resident)
    wall_s: float = 0.0          # wall-clock inside API calls (H2D + kernels + D2H)
    prior_cpu_s: float = 0.0     # CPU planar-prior stage (reported separately, like the reference's)
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    launches: int = 0
    passes: int = 0
    pass_ms: dict = field(default_factory=dict)    # stage name -> list of per-pass ms at the finest level


def pyramid_sizes(width, height, max_image_size=3200, size_bound=1000):
    """ComputeMultiScaleSettings + the per-scale size (main.cpp:35-71, :420-425)."""
    max_size = min(max(width, height), max_image_size)
    k, m = 0, max_size
    while m > size_bound:
        m //= 2
        k += 1
    return [int(max_size / (2 ** s)) for s in range(k, -1, -1)]       # coarsest first


def build_levels(scene, ref, n_src=None):
    imgs, cams, ids = scene.problem(ref)
    if n_src is not None:
        imgs, cams, ids = imgs[: n_src + 1], cams[: n_src + 1], ids[: n_src + 1]
    import cv2
    H, W = imgs[0].shape
    levels = []
    for size in pyramid_sizes(W, H):
        li, lc = synth.scale_problem(imgs, cams, size)
        nd = []
        for k, vid in enumerate(ids[1:]):
            h, w = li[k + 1].shape
            d = scene.depths_gt[vid]
            nd.append(d if d.shape == (h, w) else cv2.resize(d, (w, h), interpolation=cv2.INTER_NEAREST))
        levels.append(Level(li, lc, nd))
    return levels


def _nbytes(*arrays):
    return int(sum(a.nbytes for a in arrays if a is not None))


class B200Backend:
    """Stages through libacmmp_b200.so, GPU-resident: one context per view, the state of a stage is handed
    to the next one on the device (acmmp_next_level, acmmp_set_depth_maps with maps[0] == NULL); per level
    the host only sends the images, the neighbours' depth maps and the prior.  Results come back to the host
    where the host needs them: after the photometric stage (input of the CPU planar-prior stage) and after the
    last stage of the finest level (the view's output); `download_all=True` fetches every stage like the
    reference's RunPatchMatch does.  The context (and its buffer pool) can be kept for the next view."""
    name = "b200"

    def __init__(self, device=0, seed=1234, as_compiled=True, ctx=None, download_all=False):
        self.device, self.seed, self.as_compiled = device, seed, as_compiled
        self.ctx = ctx
        self.keep_ctx = ctx is not None
        self.download_all = download_all
        self.launches0 = ctx.launch_count() if ctx is not None else 0
        self.t = StageTimes()

    def _run(self, stage, finest, download):
        ctx = self.ctx
        download = download or self.download_all
        t0 = time.perf_counter()
        ctx.run_patch_match(download=download)
        planes = costs = None
        if download:
            planes, costs = ctx.result_host()
            self.t.d2h_bytes += _nbytes(planes, costs)
        self.t.wall_s += time.perf_counter() - t0
        tm = ctx.timings()
        self.t.gpu_ms += tm["init_ms"] + tm["pass_sum_ms"] + tm["finalize_ms"]
        self.t.passes += tm["n_pass"]
        if finest and tm["n_pass"]:
            self.t.pass_ms.setdefault(stage, []).append(tm["pass_sum_ms"] / tm["n_pass"])
        return planes, costs

    def begin_level(self, level: Level, prev=None, first=True):
        t0 = time.perf_counter()
        if self.ctx is None:
            self.ctx = Context(self.device)
        if first:
            self.ctx.set_seed(self.seed)
            self.ctx.set_plane_now_semantics(self.as_compiled)
            self.ctx.reset_modes()
            self.ctx.set_views(level.images, level.cams)
        else:
            self.ctx.next_level(level.images, level.cams)      # JBU + hierarchy hand-over on the device
            self.t.gpu_ms += self.ctx.timings()["jbu_ms"]
        self.t.wall_s += time.perf_counter() - t0
        self.t.h2d_bytes += _nbytes(*level.images)

    def photometric(self, level, finest=False):
        return self._run("photometric", finest, download=True)      # the CPU prior stage reads it

    def prior(self, level, params, masks, finest=False):
        t0 = time.perf_counter()
        self.ctx.set_planar_prior_inputs(params, masks)
        self.t.wall_s += time.perf_counter() - t0
        self.t.h2d_bytes += _nbytes(masks, params)
        return self._run("prior", finest, download=False)

    def own_depth_to(self, dev_ptr):
        """Current depth map -> a caller-owned device buffer (what a neighbour needs); waits for it."""
        self.ctx.export_depth_device(dev_ptr)
        self.ctx.synchronize()

    def geom(self, level, multi, neighbour_depths, finest=False, last=False, device_ptrs=None):
        """neighbour_depths: host arrays (uploaded here), or device_ptrs = [(ptr, w, h), ...] already on the device."""
        ctx = self.ctx
        t0 = time.perf_counter()
        ctx.reset_modes()
        ctx.set_geom_consistency(multi)
        if device_ptrs is not None:
            ctx.set_depth_maps_device([0] + [p for p, _, _ in device_ptrs], [ctx.W] + [w for _, w, _ in device_ptrs],
                                      [ctx.H] + [h for _, _, h in device_ptrs])
        else:
            ctx.set_depth_maps([None] + list(neighbour_depths))     # own map: from the device-resident state
            self.t.h2d_bytes += _nbytes(*neighbour_depths)
        self.t.wall_s += time.perf_counter() - t0
        return self._run("geom", finest, download=last)

    def end(self):
        if self.ctx is not None:
            self.t.launches += self.ctx.launch_count() - self.launches0
            if not self.keep_ctx:
                self.ctx.close()
                self.ctx = None


def run_view(levels, backend, prior_cache=None, exchange=None):
    """One reference view through every level and stage.  Returns (planes, costs) of the finest level
    -- planes = (world normal, depth) -- and leaves the timings in backend.t.
    prior_cache: dict level index -> (params, masks); filled when empty (the CPU prior stage is
    deterministic given the deterministic photometric stage, so later steps may reuse it).
    exchange: object with neighbour_depths(level_index, level, backend) -> (host_maps, device_ptrs), one of them
    None: the source views' depth maps for the next geometric stage (default: level.neighbour_depths)."""
    if prior_cache is None:
        prior_cache = {}
    state = None
    out = None
    for li, L in enumerate(levels):
        finest = li == len(levels) - 1
        backend.begin_level(L, state, first=(li == 0))
        planes, costs = backend.photometric(L, finest)
        if li not in prior_cache:
            t0 = time.perf_counter()
            dmin = float(np.float32(L.cams[0].depth_min) * np.float32(0.6))
            dmax = float(np.float32(L.cams[0].depth_max) * np.float32(1.2))
            prior_cache[li] = planar_prior(L.cams[0], np.array(planes[..., 3]), np.array(costs), dmin, dmax)
            backend.t.prior_cpu_s += time.perf_counter() - t0
        params, masks = prior_cache[li]
        planes, costs = backend.prior(L, params, masks, finest)
        for multi in (False, True):
            if exchange is None:
                host_maps, dev = L.neighbour_depths, None
            else:
                host_maps, dev = exchange.neighbour_depths(li, L, backend)
            planes, costs = backend.geom(L, multi, host_maps, finest, last=(finest and multi), device_ptrs=dev)
        if planes is not None:
            # no copy: for the B200 backend these are views of the library's pinned result buffers, valid until the
            # context runs again (the caller copies if it needs them longer); the reference backend returns fresh arrays
            out = (planes, costs)
        state = True if out is None else out        # the reference backend hands the host arrays to the next level
    return out
