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
