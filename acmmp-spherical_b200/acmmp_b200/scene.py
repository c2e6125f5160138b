"""Whole-scene schedule, GPU-resident and sharded over ranks: the reference's pipeline loop (main.cpp:417-476) for ALL views
of a scene, level by level, with the reference views dealt out to one process per GPU.

Per pyramid level (coarsest first), exactly the reference's stage order:
    sweep 1 : for every owned view: photometric stage (hierarchy + JBU hand-over above the coarsest level),
              planar prior (support points on the device, Delaunay on the host, plane fit + rasteriser on the device),
              prior stage                                                               -> what depths.dmb would hold
    exchange: the ranks all-gather the depth maps of the views they own (the path's only collective; the reference
              passes the same maps through depths.dmb files, ACMMP.cpp:653-678)
    geom 0  : for every owned view: geometric-consistency stage against its source views' maps  -> depths_geom.dmb
    exchange
    geom 1  : the same with multi_geometry
One context per owned view stays alive for the whole run (a stage finds the previous stage's planes and costs on the
device), the maps of ALL views of the level live in one device table per rank, a view's kernels read its neighbours'
maps straight from that table.  Geometric round 1 reads round 0's maps of every neighbour (Jacobi across views; the
sequential reference lets view j read the maps views i < j rewrote in the same round, main.cpp:443-445 -- a difference
inside the statistical tolerance that keeps the result independent of the number of ranks).

The C++ twin for one device is host/acmmp_main.cpp: RunResident; this module is what bench.py --config C3 and the
multi-rank tests drive.  `Worker` hides the device so that the bookkeeping (ownership, rounds, table slots, what is
exchanged when) runs under gloo on CPUs in tests/test_cpu_shard.py.
"""
from __future__ import annotations

import ctypes as C
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field

import numpy as np

from . import shard


@dataclass
class SceneTimes:
    gpu_ms: float = 0.0          # CUDA-event time of this rank's kernels
    exchange_ms: float = 0.0     # device time of the all-gathers
    exchange_bytes: int = 0
    prior_host_s: float = 0.0    # host part of the planar prior (Delaunay), overlapped with the next view's kernels
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    launches: int = 0
    passes: int = 0
    pass_ms: dict = field(default_factory=dict)


class SceneLevels:
    """Images and cameras of every view at every pyramid level, host side.  images[level][view], cams[level][view]."""

    def __init__(self, images, cams, sizes):
        self.images, self.cams, self.sizes = images, cams, sizes

    @staticmethod
    def build(scene, pin=None):
        """pin: optional callable(ndarray) -> ndarray in pinned host memory (bench.py: torch pin_memory)."""
        from . import synth
        from .pipeline import pyramid_sizes
        H, W = next(im for im in scene.images if im is not None).shape
        sizes = pyramid_sizes(W, H)
        images, cams = [], []
        for size in sizes:
            li, lc = synth.scale_problem(scene.images, scene.cams, size)
            images.append([pin(np.ascontiguousarray(im)) if pin else np.ascontiguousarray(im) for im in li])
            cams.append(lc)
        return SceneLevels(images, cams, sizes)


class GpuWorker:
    """One owned reference view on this rank's GPU: a Context kept for the whole run."""

    def __init__(self, device, seed, as_compiled=True):
        from . import Context
        self.ctx = Context(device)
        self.ctx.set_seed(seed)
        self.ctx.set_plane_now_semantics(as_compiled)
        self.first = True

    def restart(self):
        """The same view again from the coarsest level (benchmark steps): buffers and pools are kept."""
        self.first = True

    def begin_level(self, images, cams):
        if self.first:
            self.ctx.reset_modes()
            self.ctx.set_views(images, cams)
            self.first = False
            return 0.0
        self.ctx.next_level(images, cams)
        return self.ctx.timings()["jbu_ms"]

    def run(self, download=False):
        self.ctx.run_patch_match(download=download)
        return self.ctx.timings()

    def support_points(self):
        self.ctx.set_planar_prior()
        return self.ctx.support_points()

    def prior_from_triangles(self, tri_xy):
        self.ctx.planar_prior_from_triangles(tri_xy)

    def geom_mode(self, multi):
        self.ctx.reset_modes()
        self.ctx.set_geom_consistency(multi)

    def set_neighbours(self, ptrs, widths, heights):
        self.ctx.set_depth_maps_device([0] + ptrs, [self.ctx.W] + widths, [self.ctx.H] + heights)

    def export_depth(self, dev_ptr):
        self.ctx.export_depth_device(dev_ptr)

    def sync(self):
        self.ctx.synchronize()

    def result(self):
        return self.ctx.result_host()

    def launches(self):
        return self.ctx.launch_count()

    def close(self):
        self.ctx.close()


_host_lib = None


def delaunay_triangles_inside(points_xy, width, height):
    """Delaunay triangulation of the support points (host/delaunay.cpp through libacmmp_host.so; stands in for
    cv::Subdiv2D, ACMMP.cpp:932-954) -> int32 [n, 3, 2] vertices of the triangles that lie inside the image
    (main.cpp:140-146), in id order.  Releases the GIL: runs beside the next view's kernels."""
    global _host_lib
    if _host_lib is None:
        from . import PKG_DIR
        _host_lib = C.CDLL(str(PKG_DIR / "lib" / "libacmmp_host.so"))
        _host_lib.acmmp_host_delaunay_rect.restype = C.c_int
    pts = np.ascontiguousarray(points_xy, np.int32)
    n = pts.shape[0]
    if n < 3:
        return np.zeros((0, 3, 2), np.int32)
    cap = 2 * n + 16
    idx = np.zeros((cap, 3), np.int32)
    nt = _host_lib.acmmp_host_delaunay_rect(pts.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(n), C.c_int(width), C.c_int(height),
                                            idx.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(cap))
    if nt > cap:
        idx = np.zeros((nt, 3), np.int32)
        nt = _host_lib.acmmp_host_delaunay_rect(pts.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(n), C.c_int(width), C.c_int(height),
                                                idx.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(nt))
    tri = pts[idx[:nt]]                                  # [nt, 3, 2]
    inside = ((tri[..., 0] >= 0) & (tri[..., 0] < width) & (tri[..., 1] >= 0) & (tri[..., 1] < height)).all(axis=1)
    return np.ascontiguousarray(tri[inside])


class DeviceTable:
    """The depth maps of every view of one level on this rank's device.  `mine` [rounds, H, W]: the maps of the views this
    rank owns, in the order it processes them (slot k = its k-th view); `all` [world, rounds, H, W]: every rank's block after
    the all-gather, so view v = rank (v mod world)'s slot (v div world) (shard.gather_slot).
    `alloc(shape)` returns (object keeping the memory alive, device pointer of element 0)."""

    def __init__(self, world, rounds, alloc):
        self.world, self.rounds, self.alloc = world, rounds, alloc
        self.shape = None
        self.mine = self.all = None

    def fit(self, H, W):
        if self.shape != (H, W):
            self.mine, self.mine_ptr = self.alloc((self.rounds, H, W))
            self.all, self.all_ptr = self.alloc((self.world, self.rounds, H, W))
            self.shape = (H, W)

    def mine_slot_ptr(self, k):
        H, W = self.shape
        return self.mine_ptr + 4 * H * W * k

    def view_ptr(self, view):
        H, W = self.shape
        rnd, rank = shard.gather_slot(view, self.world)
        return self.all_ptr + 4 * H * W * (rank * self.rounds + rnd)


def run_scene(levels: SceneLevels, pairs, rank, world, make_worker, alloc, all_gather, finest_only_download=True,
              on_result=None, overlap_delaunay=True, workers=None, tables=None):
    """Process every view this rank owns through every level and stage.
      pairs       : [(ref view, [source views])] for every view of the scene (view id == index)
      make_worker : () -> worker object (GpuWorker, or a CPU stand-in in the tests)
      alloc       : shape -> (keep-alive, device pointer) of a float32 device array
      all_gather  : (table: DeviceTable) -> device ms; fills table.all from every rank's table.mine
      on_result   : callable(view, planes, costs) for the final result of each owned view (finest level, last stage)
      workers     : optional dict view -> worker kept from an earlier run of the same scene (their contexts keep their buffer
                    pools: nothing is allocated again); the dict is filled when empty and the workers are then left open
      tables      : optional list kept the same way for the two device tables
    Returns SceneTimes."""
    n_views = len(pairs)
    owned = shard.views_of(rank, n_views, world)
    rounds = shard.rounds(n_views, world)
    t = SceneTimes()
    keep_workers = workers is not None
    if workers is None:
        workers = {}
    for v in owned:
        if v not in workers:
            workers[v] = make_worker()
        elif hasattr(workers[v], "restart"):
            workers[v].restart()
    launches0 = {v: workers[v].launches() for v in owned}
    if tables is not None and len(tables) == 2:
        dtab, gtab = tables
    else:
        dtab, gtab = DeviceTable(world, rounds, alloc), DeviceTable(world, rounds, alloc)
        if tables is not None:
            tables[:] = [dtab, gtab]
    # the host part of a view's prior stage (the triangulation: 0.3 s for the 273 k support points of a 3200x2130 view) runs on
    # worker threads while this thread -- which issues ALL device work, in a fixed order -- goes on with the photometric stages
    # of the next views; a view is finished (prior upload, prior-stage PatchMatch, depth export) up to DEPTH views later
    DEPTH = 3
    pool = ThreadPoolExecutor(max_workers=DEPTH) if overlap_delaunay else None

    def account(tm, stage, finest):
        t.gpu_ms += tm["init_ms"] + tm["pass_sum_ms"] + tm["finalize_ms"]
        t.passes += tm["n_pass"]
        if finest and tm["n_pass"]:
            t.pass_ms.setdefault(stage, []).append(tm["pass_sum_ms"] / tm["n_pass"])

    for li in range(len(levels.sizes)):
        finest = li == len(levels.sizes) - 1
        sizes = [levels.images[li][v].shape for v in range(n_views)]
        assert len(set(sizes)) == 1, "the scene schedule needs views of one size per level"
        H, W = sizes[0]
        dtab.fit(H, W)
        gtab.fit(H, W)

        # ---- sweep 1: photometric (+ hierarchy) and prior stage
        pending = []            # (slot, view, future or triangles), oldest first

        def finish(slot, v, tri):
            w = workers[v]
            t0 = time.perf_counter()
            tri = tri.result() if hasattr(tri, "result") else tri
            t.prior_host_s += time.perf_counter() - t0            # only what was NOT hidden
            w.prior_from_triangles(tri)
            account(w.run(), "prior", finest)
            w.export_depth(dtab.mine_slot_ptr(slot))

        for slot, v in enumerate(owned):
            w = workers[v]
            ids = [v] + list(pairs[v][1])
            imgs = [levels.images[li][i] for i in ids]
            t.gpu_ms += w.begin_level(imgs, [levels.cams[li][i] for i in ids])
            t.h2d_bytes += int(sum(im.nbytes for im in imgs))
            account(w.run(), "photometric", finest)
            pts = w.support_points()
            job = pool.submit(delaunay_triangles_inside, pts, W, H) if pool else delaunay_triangles_inside(pts, W, H)
            pending.append((slot, v, job))
            while len(pending) > (DEPTH if pool else 0):
                finish(*pending.pop(0))
        while pending:
            finish(*pending.pop(0))
        for v in owned:
            workers[v].sync()
        ms = all_gather(dtab)
        t.exchange_ms += ms
        t.exchange_bytes += 4 * H * W * rounds * (world - 1)

        # ---- geometric sweeps
        for multi in (False, True):
            src_tab = gtab if multi else dtab
            last = finest and multi
            for slot, v in enumerate(owned):
                w = workers[v]
                w.geom_mode(multi)
                srcs = list(pairs[v][1])
                w.set_neighbours([src_tab.view_ptr(s) for s in srcs], [W] * len(srcs), [H] * len(srcs))
                account(w.run(download=last), "geom", finest)
                if not last:
                    # round 1 writes into the table round 1 reads from only AFTER every view of this rank has read it:
                    # round 0 exports into gtab (read by round 1), round 1 exports nothing (its maps are the output)
                    w.export_depth(gtab.mine_slot_ptr(slot))
                if last and on_result is not None:
                    planes, costs = w.result()
                    t.d2h_bytes += int(planes.nbytes + costs.nbytes)
                    on_result(v, planes, costs)
            if not multi:
                for v in owned:
                    workers[v].sync()
                ms = all_gather(gtab)
                t.exchange_ms += ms
                t.exchange_bytes += 4 * H * W * rounds * (world - 1)
    for v in owned:
        workers[v].sync()
        t.launches += workers[v].launches() - launches0[v]
        if not keep_workers:
            workers[v].close()
    if pool:
        pool.shutdown()
    return t
