"""World-size-2 gloo tests of the multi-GPU plan: views sharded per rank, depth maps all-gathered
between the geometric stages, no other collective."""
import os
import socket

import numpy as np
import pytest


def test_views_are_dealt_round_robin_and_cover_everything():
    from acmmp_b200 import shard
    for n_views, world in [(64, 8), (11, 2), (5, 4), (200, 8), (3, 1)]:
        seen = []
        for r in range(world):
            mine = shard.views_of(r, n_views, world)
            assert all(shard.owner_of(v, world) == r for v in mine)
            assert len(mine) <= shard.rounds(n_views, world)
            seen += mine
        assert sorted(seen) == list(range(n_views))
        for v in range(n_views):
            rnd, rank = shard.gather_slot(v, world)
            assert shard.views_of(rank, n_views, world)[rnd] == v


def test_neighbour_plan_prefers_fresh_maps():
    from acmmp_b200 import shard
    plan = shard.neighbour_sources([1, 2, 9, 10], round_index=0, world=8)
    assert plan == [("gathered", 1), ("gathered", 2), ("input", 9), ("input", 10)]
    plan = shard.neighbour_sources([1, 9, 17], round_index=1, world=8, computed_rounds={0})
    assert plan == [("stored", 1), ("gathered", 1), ("input", 17)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from acmmp_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W = 6, 9
        ex = shard.DepthExchange(dist, world)
        # every rank "computes" the depth map of its view of round 0: constant = 100 + view id
        view = shard.views_of(rank, 5, world)[0]
        mine = torch.full((H, W), 100.0 + view)
        gathered = ex.all_gather(mine)
        src_ids = [v for v in range(5) if v != view]
        fallbacks = [torch.full((H, W), float(v)) for v in src_ids]            # stand-ins: constant = view id
        maps = ex.pick(gathered, src_ids, 0, fallbacks)
        got = [float(m[0, 0]) for m in maps]
        want = [100.0 + v if v < world else float(v) for v in src_ids]
        ok = got == want and ex.bytes == H * W * 4 * (world - 1)
        t = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put((float(t[0]), got, want))
    finally:
        dist.destroy_process_group()


def test_depth_exchange_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, got, want = out.get(timeout=10)
    assert ok == 1.0, (got, want)


# ------------------------------------------------------------------------------------------
# the whole-scene schedule (acmmp_b200.scene.run_scene) with a CPU stand-in for the device, world size 2 over gloo
# ------------------------------------------------------------------------------------------
class _FakeWorker:
    """Stands in for scene.GpuWorker: 'depth maps' are constants encoding (view, level, stage); checks that every
    neighbour map it is handed is the one the reference's schedule says it must read."""
    log = None

    def __init__(self, view_of, n_levels):
        self.view = None
        self.level = -1
        self.stage = None
        self.view_of = view_of
        self.errors = []
        self.runs = 0

    @staticmethod
    def code(view, level, stage):
        return float(1000 * view + 10 * level + stage)

    def begin_level(self, images, cams):
        self.view = int(images[0][0, 0])          # the fake images carry their view id
        self.level += 1
        return 0.0

    def run(self, download=False):
        self.stage = {None: 0, 0: 1, 1: 2, 2: 3, 3: 0}[self.stage]        # photometric, prior, geom0, geom1, photometric ...
        self.runs += 1
        return dict(init_ms=1.0, pass_sum_ms=2.0, finalize_ms=0.5, n_pass=6)

    def support_points(self):
        return np.array([[0, 0], [4, 0], [0, 4], [4, 4]], np.int32)

    def prior_from_triangles(self, tri):
        assert tri.shape[1:] == (3, 2) and len(tri) == 2

    def geom_mode(self, multi):
        self.multi = multi

    def set_neighbours(self, ptrs, widths, heights):
        import ctypes
        want_stage = 2 if self.multi else 1       # geom 1 reads geom 0's maps, geom 0 reads the prior stage's
        got = [ctypes.cast(p, ctypes.POINTER(ctypes.c_float))[0] for p in ptrs]
        want = [self.code(s, self.level, want_stage) for s in self.view_of[self.view]]
        if got != want:
            self.errors.append((self.view, self.level, self.multi, got, want))
        _FakeWorker.log.append(self)

    def export_depth(self, dev_ptr):
        import ctypes
        ctypes.cast(dev_ptr, ctypes.POINTER(ctypes.c_float))[0] = self.code(self.view, self.level, self.stage)

    def sync(self):
        pass

    def result(self):
        return np.zeros((1, 1, 4), np.float32), np.zeros((1, 1), np.float32)

    def launches(self):
        return 7 * self.runs

    def close(self):
        pass


def _scene_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from acmmp_b200 import scene
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_views, n_levels = 5, 2
        pairs = [(v, [(v + 1) % n_views, (v + 2) % n_views, (v - 1) % n_views]) for v in range(n_views)]
        view_of = {v: srcs for v, srcs in pairs}
        images = [[np.full((6 * (l + 1), 8 * (l + 1)), float(v), np.float32) for v in range(n_views)] for l in range(n_levels)]
        levels = scene.SceneLevels(images, [[None] * n_views for _ in range(n_levels)], [8, 16])
        _FakeWorker.log = []
        keep = []

        def alloc(shape):
            t = torch.zeros(shape, dtype=torch.float32)
            keep.append(t)
            return t, t.data_ptr()

        def all_gather(table):
            parts = [torch.empty_like(table.mine) for _ in range(world)]
            dist.all_gather(parts, table.mine)
            for k, p in enumerate(parts):
                table.all[k].copy_(p)
            return 0.25
        results = []
        t = scene.run_scene(levels, pairs, rank, world, lambda: _FakeWorker(view_of, n_levels), alloc, all_gather,
                            on_result=lambda v, p, c: results.append(v), overlap_delaunay=(rank == 0))
        errors = [e for w in _FakeWorker.log for e in w.errors]
        mine = [v for v in range(n_views) if v % world == rank]
        ok = (not errors and sorted(results) == mine and t.passes == 6 * 4 * n_levels * len(mine)
              and abs(t.exchange_ms - 0.25 * 2 * n_levels) < 1e-9 and t.launches == 7 * 4 * n_levels * len(mine))
        flag = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put((float(flag[0]), errors[:3], results, t.passes))
    finally:
        dist.destroy_process_group()


def test_whole_scene_schedule_over_gloo_world_size_2():
    """Every geometric stage of every view reads exactly the maps the reference's schedule prescribes (the prior stage's
    maps of the same level in round 0, round 0's maps in round 1), whichever rank computed them."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scene_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, errors, results, passes = out.get(timeout=10)
    assert ok == 1.0, (errors, results, passes)
